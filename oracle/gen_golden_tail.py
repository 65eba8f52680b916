"""Golden vectors for the encoder tail from the reference's OWN class source.

`models/skip_vid_generator/models/skip_autoencoder.py` cannot be imported here (its package __init__ JIT-compiles CUDA
extensions and imports cupy; the module itself imports torchvision), so this script reads the file, extracts the source
of `EqualConv2d` and `ConvLayer` with `ast`, and executes exactly those class definitions (unmodified) in a namespace
that provides what they reference (torch, nn, F, math; `Blur` is never instantiated for kernel_size = 1 without
up/down-sampling).  The fixtures hold inputs, the block's parameters and the outputs of the reference classes on CPU.

    python oracle/gen_golden_tail.py          (needs /root/reference; writes tests/golden/tail_*.npz)
"""
import ast
import math
import os
import sys

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

REF = "/root/reference/models/skip_vid_generator/models/skip_autoencoder.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_classes():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "F": F, "math": math, "Blur": None}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("EqualConv2d", "ConvLayer"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns["EqualConv2d"], ns["ConvLayer"]


def main():
    _, ConvLayer = reference_classes()
    cases = {
        "tail_c128_o64": (3, 128, 64, 8, 8, 1),          # G, C_in, C_out, h, w, seed
        "tail_c512_o256": (2, 512, 256, 8, 8, 2),        # BAIR default: block_out 512 -> z_size 256, 8x8 latents
        "tail_c64_o512_ragged": (1, 64, 512, 5, 7, 3),   # two output-channel tiles, ragged position count
    }
    for name, (G, ci, co, h, w, seed) in cases.items():
        torch.manual_seed(seed)
        layer = ConvLayer(ci, co, 1)                     # the reference block, parameters as it initialises them
        with torch.no_grad():
            layer[0].bias.copy_(torch.randn(co) * 0.5)   # (zeros at init: give the bias something to do)
        x = torch.randn(G, ci, h, w) * 1.5
        with torch.no_grad():
            out = layer(x)
            out_n = out / torch.norm(out, p=2, dim=1, keepdim=True)      # skip_autoencoder.py:348-349
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), weight=layer[0].weight.detach().numpy(),
                            bias=layer[0].bias.detach().numpy(), out=out.numpy(), out_normalized=out_n.numpy())
        print(name, tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    sys.exit(main())
