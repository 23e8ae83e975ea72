#!/bin/bash
# usage: tools/gpu_round.sh <tag> [bench args...]   -- tests, bench, ncu launch list + full capture of our kernels
tag=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$tag.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)[0-9_]" -s 6 -c 2 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_f_$tag.log 2>&1
cat gpurun_out/tests_$tag.log gpurun_out/bench_$tag.json
tail -3 gpurun_out/bench_$tag.err
