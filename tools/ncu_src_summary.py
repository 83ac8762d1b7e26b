"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS lines."""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
def num(x):
    try: return float(x)
    except ValueError: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(num(r[col[s]]) for r in data) for s in stalls}
allsmp = sum(num(r[col["# Samples"]]) for r in data)
inst = sum(num(r[col["Instructions Executed"]]) for r in data)
tinst = sum(num(r[col["Thread Instructions Executed"]]) for r in data)
print(f"SASS lines {len(data)}  samples {allsmp:.0f}  warp-instr {inst:.3g}  threads/instr {tinst / max(inst, 1):.1f}")
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {s:28s} {100 * v / max(allsmp, 1):5.1f}%")
print("hottest lines:")
for r in sorted(data, key=lambda r: -num(r[col["# Samples"]]))[:top]:
    best = max(stalls, key=lambda s: num(r[col[s]]))
    print(f"  {r[col['Address']][-5:]} {num(r[col['# Samples']]):7.0f} {num(r[col['Instructions Executed']]):10.0f} {num(r[col['Avg. Threads Executed']]):5.1f} {best[6:]:14s} {r[col['Source']][:90]}")
