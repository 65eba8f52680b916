"""Compact view of bench.py JSON lines.  usage: python tools/show_bench.py file.json [...]"""
import json, sys
for p in [a for a in sys.argv[1:] if a != "-v"]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f"{p} unreadable: {e}")
        continue
    r = d.get("roofline") or {}
    c = d.get("clocks") or {}
    print(f"{p}: {d['config']['workload']} {d['value'] / 1e6:.1f} M/s {d['ms_per_step']:.4f} ms/step | screen {r.get('achieved', 0):.0f} TF/s "
          f"({100 * (r.get('frac') or 0):.1f}%) {r.get('ms_per_launch', 0):.4f} ms | e2e {d['e2e']['value'] / 1e6:.1f} M/s | "
          f"launches {d.get('gpu_launches')} host {d.get('host_issue_ms_per_step', 0):.3f} ms | sm {c.get('sm_mhz')} {c.get('reasons')}")
    if "-v" in sys.argv:
        continue
    print("    ", {k.replace('ccvsq_', ''): round(v['ms_per_step'], 4) for k, v in d.get("kernel_breakdown", {}).items()})
    print("    ", {k: (round(v['achieved_GBs']), round(v['frac_of_hbm_peak'], 3)) for k, v in d.get("hbm_kernels", {}).items()})
