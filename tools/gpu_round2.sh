#!/bin/bash
# round-2 evidence run on one B200: bench lines of the named configs, the wide-window line, ncu launch list and full
# captures of the dominant kernels (reference numerics and closed form).  Outputs under gpurun_out/r2_*.
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; tail -1 gpurun_out/r2_bench_$tag.err; }
run cfg5
run cfg4 --workload cfg4
run cfg3 --workload cfg3 --no-cpu-baseline
run cfg2 --workload cfg2 --steps 1000 --no-cpu-baseline
run cfg1 --workload cfg1 --steps 1000 --no-cpu-baseline
run cfg5_wide --p-window 2000:5000 --no-cpu-baseline
# launch list of the default command (few steps), then full captures of the two pairs on the same grid (K = 2)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/r2_launches_cfg5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)4" -s 2 -c 2 -o gpurun_out/r2_prof_dg4 -f \
    python tools/prof_step.py cfg5 reference 2 3 > gpurun_out/r2_ncu_f_dg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)_cf2" -s 2 -c 2 -o gpurun_out/r2_prof_cf2 -f \
    python tools/prof_step.py cfg5 closed_form 2 3 > gpurun_out/r2_ncu_f_cf2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)_gc2" -s 2 -c 2 -o gpurun_out/r2_prof_gc2 -f \
    python tools/prof_step.py cfg4 reference 4 3 > gpurun_out/r2_ncu_f_gc2.log 2>&1
ls -la gpurun_out/r2_* | tail -30
