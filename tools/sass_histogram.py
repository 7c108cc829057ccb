"""SASS evidence per kernel of libtgfr_b200.so: counts of the tcgen05 / TMA / mbarrier mnemonics (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMAREDG / UBLKCP = TMA tensor loads / reduce-adds / bulk copies, SYNCS =
mbarrier ops, MUFU, FFMA, HFMA2 ...) -> profiles/r2_sass_histogram.txt.   python tools/sass_histogram.py > profiles/..."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "text_guided_face_recognition_b200", "libtgfr_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMAREDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "USETMAXREG",
        "MUFU", "FFMA", "HFMA2", "F2FP", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "BAR", "NANOSLEEP")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
hist, cur, total = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("tgfr::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\((?!anonymous).*", "", name)
        cur = hist.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                cur[k] += 1
                total[k] += 1
print("# cuobjdump -sass libtgfr_b200.so: instruction counts per kernel (static code, not executions)")
print("# totals: " + ", ".join(f"{k} {v}" for k, v in total.items() if v))
for name, c in hist.items():
    tags = ", ".join(f"{k} {c[k]}" for k in KEYS if c[k])
    print(f"{c['_n']:6d} instr  {name[:110]:110s} {tags}")
