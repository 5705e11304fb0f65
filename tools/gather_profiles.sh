#!/bin/bash
# Round-end evidence on ONE box: GPU tests, parity table, full bench line, per-launch table, step timeline,
# ncu launch list (gpu__time_duration + DRAM bytes per launch) of the same bench command.
tag=${1:-r02}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo smoke_rc=$?; tail -1 gpurun_out/${tag}_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo pytest_rc=$?
python tools/parity_table.py --out gpurun_out/${tag}_parity_table.md > gpurun_out/${tag}_parity.log 2>&1; echo parity_rc=$?
python bench.py --dump-launches gpurun_out/${tag}_launches_step.txt > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo bench_rc=$?
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo ref_rc=$?
python tools/timeline.py --out gpurun_out/${tag}_timeline.txt > /dev/null 2> gpurun_out/${tag}_timeline.err; echo timeline_rc=$?
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/${tag}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-stock --no-cpu-baseline --no-also \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo ncu_rc=$?
python tools/agg_launches.py gpurun_out/${tag}_ncu_launches.csv > gpurun_out/${tag}_ncu_launches_by_kernel.txt 2>&1; tail -3 gpurun_out/${tag}_ncu_launches_by_kernel.txt
tail -3 gpurun_out/${tag}_pytest.log
