"""Diagnostic: does clock sampling (nvidia-smi -lms / in-process NVML) stall kernel launches?"""
import sys, time, os, subprocess, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops, VectorQuantizer

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs(wl, dev, 1234)
vq = VectorQuantizer(K, D, 0.25).to(dev).eval()
with torch.no_grad():
    vq.embedding.weight.copy_(cb)


def step():
    with torch.no_grad():
        z_q, loss, (perp, _, idx) = vq(z)
        dec = vq.embed_code(idx.view(clips * frames, h, w_))
    return idx, loss, perp, dec


def run(tag, reps=10, steps=20):
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    hosts, devs = [], []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = step()
        e1.record()
        hosts.append((time.perf_counter() - t0) / steps * 1e3)
        torch.cuda.synchronize()
        devs.append(e0.elapsed_time(e1) / steps)
        time.sleep(0.03)
    print(f"{tag:28s} device ms/step min {min(devs):.3f} med {sorted(devs)[len(devs)//2]:.3f} max {max(devs):.3f} | "
          f"host ms/step min {min(hosts):.3f} max {max(hosts):.3f}", flush=True)


run("no sampler")
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=" + bench.ClockSampler.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.5)
run("nvidia-smi -lms 100")
p.terminate(); p.wait()
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.5)
run("nvidia-smi small query")
p.terminate(); p.wait()

import pynvml
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = False
samples = []
def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)))
        time.sleep(0.01)
t = threading.Thread(target=poll); t.start()
run("pynvml thread 10 ms")
stop = True; t.join()
print(len(samples), "nvml samples; last", samples[-1], "max sm", pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM))
run("no sampler again")
