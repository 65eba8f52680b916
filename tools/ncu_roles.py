"""Per-role stall summary of the screen kernel from an `ncu --page source --csv` export.
usage: python tools/ncu_roles.py file.src.csv b0 b1 b2 ... (line boundaries between roles)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}; data = rows[hi + 1:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
bounds = [int(x) for x in sys.argv[2:]] + [len(data)]
a = 0
for b in bounds:
    tot = 0; agg = {}; inst = 0
    for i in range(a, b):
        r = data[i]; tot += float(r[col['# Samples']]); inst += float(r[col['Instructions Executed']])
        for h in stall_cols: agg[h] = agg.get(h, 0) + float(r[col[h]])
    print(f"lines {a:5d}-{b:5d}: samples {int(tot):8d} inst {inst:.3e}", {k[6:]: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.03})
    a = b
