"""Aggregates an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch
list of bench.py into per-kernel totals for ONE trainer step (the launches between two consecutive
rmsprop_conv_packed launches) and the DRAM traffic of the convolution classes.

    python tools/agg_launches.py profiles/r02_ncu_launches_step.csv > profiles/r02_ncu_launches_by_kernel.txt
"""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
data, names, order = collections.defaultdict(dict), {}, []
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    x = dict(zip(hdr, r))
    i = int(x["ID"])
    if i not in names:
        order.append(i)
    names[i] = x["Kernel Name"]
    data[i][x["Metric Name"]] = (float(x["Metric Value"].replace(",", "")), x["Metric Unit"])
to_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
ends = [k for k, i in enumerate(order) if "rmsprop_conv_packed" in names[i]]
if len(ends) >= 2:
    sel = order[ends[-2] + 1: ends[-1] + 1]
    what = f"one step = launches {sel[0]}..{sel[-1]} (between two rmsprop_conv_packed launches)"
else:
    sel, what = order, "all profiled launches (no full step delimited)"
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i in sel:
    m = data[i]
    n = names[i].split("(")[0].replace("void ", "").replace("ugp::", "")
    t = m["gpu__time_duration.sum"][0] * to_us[m["gpu__time_duration.sum"][1]]
    b = sum(m[k][0] * to_b[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
    agg[n][0] += 1
    agg[n][1] += t
    agg[n][2] += b
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {what}; {len(sel)} launches, {tot:.1f} us serialised under ncu (cold caches, --clock-control none)")
print("#      us   share   n   DRAM MB  kernel")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} {100 * v[1] / tot:6.1f}% {v[0]:4d} {v[2] / 1e6:9.1f}  {n[:90]}")
fd = [v for n, v in agg.items() if n.startswith(("conv_gemm_kernel", "conv_pair_kernel", "conv_halo_kernel", "conv_halo_pair_kernel"))]
wg = [v for n, v in agg.items() if n.startswith("conv_wgrad")]
print(f"# conv forward + dgrad class: {sum(v[0] for v in fd)} launches, {sum(v[1] for v in fd):.1f} us, "
      f"DRAM {sum(v[2] for v in fd) / 1e6:.1f} MB per step")
print(f"# conv wgrad class: {sum(v[0] for v in wg)} launches, {sum(v[1] for v in wg):.1f} us, "
      f"DRAM {sum(v[2] for v in wg) / 1e6:.1f} MB per step")
fl = [v for n, v in agg.items() if n.startswith(("conv_first_", "im2col_first"))]
print(f"# first layer (fused row building, HBM-bound; outside both classes): {sum(v[0] for v in fl)} launches, "
      f"{sum(v[1] for v in fl):.1f} us, DRAM {sum(v[2] for v in fl) / 1e6:.1f} MB per step")
