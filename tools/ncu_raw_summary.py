"""`ncu -i rep --page raw --csv` (stdin) -> the "== kernel / metric = value" text kept under profiles/ (one block per
kernel launch; `--first` keeps the first launch of every distinct kernel name).  bench.py reads dram bytes from it.

    ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv | python tools/ncu_raw_summary.py --first > profiles/r2_final_ncu_raw.txt
"""
import csv
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]
UNIT_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
UNIT_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    first = "--first" in sys.argv
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    seen = set()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        if first and name in seen:
            continue
        seen.add(name)
        print("== " + name[:150])
        for k in KEEP:
            if k not in ix or r[ix[k]] in ("", "n/a"):
                continue
            v = float(r[ix[k]].replace(",", ""))
            u = units[ix[k]]
            if k.startswith("dram__bytes"):
                v *= UNIT_MB.get(u, 1.0)
            if k == "gpu__time_duration.sum":
                v *= UNIT_US.get(u, 1.0)
            print(f"{k} = {v:.6f}" if isinstance(v, float) and not k.startswith("launch__") else f"{k} = {v:g}")
        st = []
        for c in stall_cols:
            if r[ix[c]] not in ("", "n/a"):
                st.append((float(r[ix[c]].replace(",", "")), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        st.sort(reverse=True)
        if st:
            print("stalls = " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]))
        print()


if __name__ == "__main__":
    main()
