"""Summarise an `ncu --page source --csv` export: top instructions by executed count / stall samples.
usage: python tools/ncu_hot.py file.src.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
def f(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print(f"total warp instructions {tot_inst:.3e}, samples {tot_samp:.0f}, sass lines {len(data)}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_st = {h: sum(f(r, h) for r in data) for h in stall_cols}
print("stall totals:", {k[6:]: int(v) for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1]) if v > 0})
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
print(f"\n-- address order, lines with >={thr*100}% of instructions or samples")
for i, r in enumerate(data):
    ie, s = f(r, "Instructions Executed"), f(r, "# Samples")
    if ie >= thr * tot_inst or s >= thr * tot_samp:
        st = sorted(((f(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {ie/tot_inst*100:5.2f}%i {s/tot_samp*100:5.2f}%s  {r[col['Source']].strip()[:70]:70s} {st[0][1]}:{int(st[0][0])} {st[1][1]}:{int(st[1][0])}")
