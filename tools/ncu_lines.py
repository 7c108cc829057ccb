"""Stall samples per SOURCE LINE: joins `ncu -i rep --page source --csv` (per SASS instruction, absolute addresses) with
`nvdisasm -g -c` of the same cubin (per instruction offsets + line info).

    python tools/ncu_lines.py <source.csv> <nvdisasm.txt> <mangled-kernel-substring> [ntop]
"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[k]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[k + 1:] if len(r) == len(hdr)]
base = min(int(r[ix['Address']], 16) for r in data)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
# offsets -> line
off2line = {}
cur = None; active = False
for line in open(sys.argv[2]):
    if line.startswith('//---------------------'):
        active = sys.argv[3] in line
        continue
    if not active:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', line)
    if m:
        off2line[int(m.group(1), 16)] = cur
agg = {}
tot = 0
for r in data:
    n = int(r[ix['# Samples']] or 0)
    if not n:
        continue
    tot += n
    ln = off2line.get(int(r[ix['Address']], 16) - base, ("?", -1))
    a = agg.setdefault(ln, [0, {}])
    a[0] += n
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            a[1][c[6:]] = a[1].get(c[6:], 0) + v
_files = {}


def src_line(f, ln):
    if f not in _files:
        try:
            _files[f] = open(f).read().splitlines()
        except OSError:
            _files[f] = []
    L = _files[f]
    return L[ln - 1].strip()[:80] if 0 < ln <= len(L) else '?'


ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
print("total samples", tot)
for ln, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    text = src_line(*ln)
    print(f"{n:6d} {100*n/tot:5.1f}% {ln[0].split('/')[-1][:12]:12s}:{ln[1]:<5d} {text:80s} {top}")
