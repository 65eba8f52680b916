"""Print a compact summary of bench.py JSON lines (one file per argument)."""
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable:", e); continue
    r = d.get("roofline") or {}
    print(f"{d['config']['workload']}: {d['value']/1e6:.1f} M latents/s  {d['ms_per_step']:.4f} ms/step  "
          f"screen {r.get('achieved') or 0:.0f} TF/s ({(r.get('frac') or 0)*100:.1f}% of burst)  e2e {d['e2e']['value']/1e6:.1f} M/s  "
          f"launches {d.get('gpu_launches')}  host {d.get('host_issue_ms_per_step', 0):.3f} ms  clocks {d.get('clocks')}")
    print("   ", {k.replace('ccvsq_', ''): round(v["ms_per_step"], 4) for k, v in d.get("kernel_breakdown", {}).items()})
    print("   ", d.get("hbm_kernels"))
