"""One line per profiled launch of an `ncu --set full` report: duration, DRAM bytes and throughput, issue
utilisation, the busiest pipe - the numbers DESIGN.md quotes for the HBM-bound kernels.

    ncu -i gpurun_out/r02h_hbm.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/ncu_summary.py /tmp/raw.csv
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
col = {n: i for i, n in enumerate(hdr)}


def get(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default


want = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"),
        ("sm__instruction_throughput.avg.pct_of_peak_sustained_active", "inst_%"),
        ("smsp__issue_active.avg.pct", "issue_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
print("# " + sys.argv[1])
print("# kernel | " + " | ".join(f"{short} [{units[col[name]]}]" for name, short in want if name in col))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = get(r, "Kernel Name").split("(")[0].replace("void ", "").replace("ugp::", "")
    print(f"{name[:52]:52s} | " + " | ".join(get(r, n) for n, _ in want if n in col))
