#!/bin/bash
# session-3 evidence run on one B200: GPU tests, bench lines of the named configs, ncu launch list + full captures.
# Outputs under gpurun_out/r3_*.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r3_gpu_tests.log; tail -3 gpurun_out/r3_gpu_tests.log
run() { tag=$1; shift; timeout 900 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/r3_bench_$tag.json 2> gpurun_out/r3_bench_$tag.err; tail -1 gpurun_out/r3_bench_$tag.err; }
run cfg5
run cfg4 --workload cfg4 --no-cpu-baseline
run cfg3 --workload cfg3 --no-cpu-baseline
run cfg2 --workload cfg2 --steps 1000 --no-cpu-baseline
run cfg5_wide --p-window 2000:5000 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/r3_launches_cfg5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_l.log 2>&1
ls -la gpurun_out/r3_* | tail -30
