"""Hot SASS instructions of one kernel from `ncu --page source --csv`:  python tools/ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[idx["# Samples"]]) for r in data)
totinst = sum(int(r[idx["Instructions Executed"]]) for r in data)
print(f"total samples {tot}, warp instructions executed {totinst}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]]) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[idx[h]]), h) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[idx['# Samples']]):6d} {int(r[idx['Instructions Executed']]):8d}  {r[idx['Source']].strip()[:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
