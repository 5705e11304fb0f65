#!/bin/bash
# A/B of two builds of libugp_b200.so on one box (alternating bench runs): build the old revision, copy its
# library to ab/libugp_prev.so (git-ignored, travels with gpurun), then `gpurun -- bash tools/ab_bench.sh`.
for i in 1 2 3; do
  for v in prev new; do
    if [ $v = prev ]; then export UGP_LIB_PATH=ab/libugp_prev.so; else unset UGP_LIB_PATH; fi
    python bench.py --steps 40 --warmup 8 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$v', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1))"
  done
done
