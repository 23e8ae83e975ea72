#!/bin/bash
# fifth (final) evidence run on one B200: GPU tests, bench lines of the named configs, ncu launch list, full captures of
# the three kernel pairs, DRAM traffic at the bench's batch size.  Outputs under gpurun_out/r5_*.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -70 > gpurun_out/r5_gpu_tests.log; tail -2 gpurun_out/r5_gpu_tests.log
run() { tag=$1; shift; timeout 900 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/r5_bench_$tag.json 2> gpurun_out/r5_bench_$tag.err; tail -1 gpurun_out/r5_bench_$tag.err; }
run cfg5
run cfg4 --workload cfg4 --no-cpu-baseline
run cfg3 --workload cfg3 --no-cpu-baseline
run cfg2 --workload cfg2 --steps 1000 --no-cpu-baseline
run cfg1 --workload cfg1 --steps 1000 --no-cpu-baseline
run cfg5_wide --p-window 2000:5000 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/r5_launches_cfg5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)4" -s 2 -c 2 -o gpurun_out/r5_prof_dg4 -f \
    python tools/prof_step.py cfg5 reference 2 3 > gpurun_out/r5_ncu_f_dg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)_cf2" -s 2 -c 2 -o gpurun_out/r5_prof_cf2 -f \
    python tools/prof_step.py cfg5 closed_form 2 3 > gpurun_out/r5_ncu_f_cf2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)_gc2" -s 2 -c 2 -o gpurun_out/r5_prof_gc2 -f \
    python tools/prof_step.py cfg4 reference 4 3 > gpurun_out/r5_ncu_f_gc2.log 2>&1
for nm in "reference:k_(fwd|adj)4:dg4" "closed_form:k_(fwd|adj)_cf2:cf2"; do
  IFS=: read num rx tag <<< "$nm"
  timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"$rx" -s 2 -c 2 -o gpurun_out/r5_traffic_$tag -f python tools/prof_step.py cfg5 $num 8 3 > gpurun_out/r5_traffic_$tag.log 2>&1
done
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_(fwd|adj)_gc2" -s 2 -c 2 -o gpurun_out/r5_traffic_gc2 -f python tools/prof_step.py cfg4 reference 32 3 > gpurun_out/r5_traffic_gc2.log 2>&1
ls -la gpurun_out/r5_* | tail -40
