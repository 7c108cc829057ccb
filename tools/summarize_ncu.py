"""`ncu -i <rep> --page raw --csv` -> one line per kernel launch with the metrics the design discussion uses.
    ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv | python tools/summarize_ncu.py > profiles/<name>.txt"""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us", 1e-3 if units[ix["gpu__time_duration.sum"]] == "ns" else 1.0),
        ("dram__bytes_read.sum", "rdMB", None), ("dram__bytes_write.sum", "wrMB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("sm__warps_active.avg.per_cycle_active", "warps", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("launch__grid_size", "grid", 1.0), ("launch__block_size", "block", 1.0)]


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


print("# " + " ".join(f"{n:>8s}" for _, n, _ in cols) + "  kernel")
for r in rows[2:]:
    out = []
    for key, _, scale in cols:
        if key not in ix or r[ix[key]] in ("", "n/a"):
            out.append(f"{'-':>8s}")
            continue
        v = to_mb(r[ix[key]], units[ix[key]]) if scale is None else float(r[ix[key]].replace(",", "")) * scale
        out.append(f"{v:8.1f}")
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("tgfr::<unnamed>::", "").replace("void ", "")
    print("  " + " ".join(out) + "  " + name[:90])
