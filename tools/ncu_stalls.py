"""Top stalled SASS instructions of a kernel from `ncu -i rep --page source --csv` (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[k]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[k + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:ntop]:
    n = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{n:6d} {100*n/tot:5.1f}% {r[ix['Address']][-5:]} {r[ix['Source']][:76]:76s} {st}")
agg = {c[6:]: sum(int(r[ix[c]] or 0) for r in data) for c in stall_cols}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
