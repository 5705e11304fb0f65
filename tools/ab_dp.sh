#!/bin/bash
# Data-parallel A/B on one box: tools/ab_dp.sh NGPUS tag "ENV=V ENV=V" "ENV=V" ...   ("" = defaults)
n=$1; tag=$2; shift; shift
mkdir -p gpurun_out
i=0
for e in "$@"; do
  env $e python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + i)) \
      bench.py --gpus $n --steps 30 --no-stock --no-cpu-baseline ${BENCH_EXTRA:---no-also} \
      > gpurun_out/${tag}_dp${n}_$i.json 2> gpurun_out/${tag}_dp${n}_$i.err
  rc=$?
  python - "$e" gpurun_out/${tag}_dp${n}_$i.json $rc <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]).read().strip().splitlines() if l.startswith("{")][-1])
    also = {k: round(v["ms_per_step"], 3) for k, v in (d.get("also") or {}).items()}
    print(f"[{sys.argv[1]:44s}] rc={sys.argv[3]} n={d['n_gpus']} {d['ms_per_step']:.3f} ms/step {d['value']:.0f} img/s  e2e {d['e2e']['value']:.0f}  {also}", flush=True)
except Exception as ex:
    print(f"[{sys.argv[1]}] rc={sys.argv[3]} no line ({ex})", flush=True)
PY
  tail -2 gpurun_out/${tag}_dp${n}_$i.err | cut -c1-300
  i=$((i+1))
done
