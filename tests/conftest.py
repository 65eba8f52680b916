"""pytest configuration: the `gpu` marker, golden-fixture loading and import paths.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI load/export checks (no GPU work).
`-m gpu`       : parity of the CUDA path (through the C ABI) against the oracle and the fixtures.
"""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
_ALL_NPZ = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GOLDEN_NAMES = [n for n in _ALL_NPZ if not n.startswith("tail_")]        # quantizer fixtures (oracle/gen_golden.py)
TAIL_NAMES = [n for n in _ALL_NPZ if n.startswith("tail_")]              # encoder-tail fixtures (oracle/gen_golden_tail.py)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One fixture produced by oracle/gen_golden.py from the unmodified reference."""

    def __init__(self, name):
        self.name = name
        d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.raw = d
        t = lambda k: torch.from_numpy(d[k])
        self.z, self.codebook = t("z"), t("codebook")
        self.beta = float(d["beta"])
        self.mult = int(d["mult"])
        self.normalize = bool(d["normalize"])
        self.n_e = int(d["n_e"])
        self.e_dim_total = int(d["e_dim_total"])
        self.z_q, self.loss, self.perplexity = t("z_q"), t("loss"), t("perplexity")
        self.indices = t("indices")
        self.one_hot_sum = t("one_hot_sum")
        self.g_zq, self.g_loss = t("g_zq"), float(d["g_loss"])
        self.dz, self.dE = t("dz"), t("dE")
        self.top2_rel_gap = t("top2_rel_gap")
        self.code = t("code") if "code" in d.files else None
        self.embedded = t("embedded") if "embedded" in d.files else None

    @property
    def e_dim(self):
        return self.e_dim_total // self.mult


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return Golden(request.param)


class TailGolden:
    """One encoder-tail fixture produced by oracle/gen_golden_tail.py from the reference's own ConvLayer source."""

    def __init__(self, name):
        self.name = name
        d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        t = lambda k: torch.from_numpy(d[k])
        self.x, self.weight, self.bias, self.out, self.out_normalized = t("x"), t("weight"), t("bias"), t("out"), t("out_normalized")


@pytest.fixture(params=TAIL_NAMES)
def tail_golden(request):
    return TailGolden(request.param)
