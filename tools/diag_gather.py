"""Why is the decode gather slower back to back than alone?  Per-launch CUDA-event timing of the row-major gather
(and yardsticks) in several launch patterns, with the SM clock sampled through NVML while each pattern runs.
usage: python tools/diag_gather.py [workload]"""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs(wl, dev, 1234)
lay = ops.layout_of(z.shape, D, 1)
w = cb.contiguous()
idx = ops.quantize_forward(z, lay, w, 0.25, indices_only=True).idx
del z
out = torch.empty(n, D, device=dev)
out2 = torch.empty(n, D, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
rl = ops.rows_layout(n, D)

import pynvml
pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)


class Clk:
    def __enter__(self):
        self.s, self.p, self.stop = [], [], False
        def run():
            while not self.stop:
                self.s.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
                self.p.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0)
                time.sleep(0.002)
        self.t = threading.Thread(target=run); self.t.start(); return self
    def __exit__(self, *a):
        self.stop = True; self.t.join()
    def txt(self):
        s = sorted(self.s)
        return f"sm clock median {s[len(s)//2]} MHz (min {s[0]}, max {s[-1]}), power max {max(self.p):.0f} W, {len(s)} samples"


def g(o=out):
    ops._call("ccvsq_gather", ops._ptr(idx), ops._ptr(w), K, rl, ops._ptr(o), ops._ptr(err), ops._stream(dev))


def pattern(name, fn, reps=200, gap=None):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    with Clk() as c:
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for a, b in evs:
            if gap:
                torch.cuda._sleep(gap)
            a.record(); fn(); b.record()
        t1.record()
        torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    tot = t0.elapsed_time(t1) / reps
    gb = n * (4 * D + 8) / 1e9
    print(f"{name:46s} per-launch median {ts[len(ts)//2]*1e3:8.1f} us (best {ts[0]*1e3:8.1f}) = {gb/ts[len(ts)//2]*1e3:6.0f} GB/s | "
          f"loop {tot*1e3:8.1f} us/iter | {c.txt()}", flush=True)


print(f"workload {wl}: {n} rows x {D}, K={K}; PDL {'off' if os.environ.get('CCVSQ_NO_PDL') else 'on'}")
pattern("gather rows, back to back, same buffer", g)
pattern("gather rows, alternating two buffers", lambda s=[0]: (g(out if s[0] % 2 == 0 else out2), s.__setitem__(0, s[0] + 1)))
pattern("gather rows, 1 ms idle gap between launches", g, reps=50, gap=2_000_000)
pattern("gather rows via ops.gather (fresh alloc)", lambda: ops.gather(idx, w))
pattern("torch fill, back to back", lambda: out.fill_(1.0))
pattern("torch fill, 1 ms idle gap", lambda: out.fill_(1.0), reps=50, gap=2_000_000)
pattern("torch copy out2<-out, back to back", lambda: out2.copy_(out))
idx0 = torch.zeros_like(idx)
def g0():
    ops._call("ccvsq_gather", ops._ptr(idx0), ops._ptr(w), K, rl, ops._ptr(out), ops._ptr(err), ops._stream(dev))
pattern("gather rows, all codes = 0, back to back", g0)
lay_cm = lay
outc = out.view(-1)
def gc():
    ops._call("ccvsq_gather", ops._ptr(idx), ops._ptr(w), K, lay_cm, ops._ptr(outc), ops._ptr(err), ops._stream(dev))
pattern("gather channel-major, back to back", gc)
pattern("gather channel-major, 1 ms idle gap", gc, reps=50, gap=2_000_000)
