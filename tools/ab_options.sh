#!/bin/bash
# A/B of ugp_set_option switches on ONE box: tools/ab_options.sh tag "opt=v,opt=v" "opt=v" ...
# Each configuration runs bench.py (device-timed step + per-launch dump) with UGP_OPTIONS set.
tag=$1; shift
mkdir -p gpurun_out
i=0
for o in "$@"; do
  UGP_OPTIONS="$o" python bench.py --no-stock --no-cpu-baseline --no-also --steps 30 \
      --dump-launches gpurun_out/${tag}_launches_$i.txt > gpurun_out/${tag}_bench_$i.json 2> gpurun_out/${tag}_bench_$i.err
  python - "$o" gpurun_out/${tag}_bench_$i.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
h = d["roofline"]["hbm"]["kernels"]
print(f"[{sys.argv[1]:28s}] {d['ms_per_step']:.3f} ms/step  kernel sum {d['roofline']['kernel_time_sum_ms']:.3f} ms  "
      + "  ".join(f"{k}={v['us_per_step']:.0f}" for k, v in h.items() if k.startswith(("bn_", "upsample"))), flush=True)
PY
  i=$((i+1))
done
