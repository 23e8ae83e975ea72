#!/bin/bash
# usage: tools/gpu_prof.sh <tag> <kernel regex> <workload> <numerics> [K]   -- launch list + ncu --set full of the matching kernels
tag=$1; rx=$2; wl=$3; nm=$4; K=${5:-2}
mkdir -p gpurun_out
python tools/prof_step.py $wl $nm $K 3 > gpurun_out/prof_${tag}.log 2>&1 || { tail -5 gpurun_out/prof_${tag}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${tag}.csv \
    python tools/prof_step.py $wl $nm $K 3 > gpurun_out/ncu_l_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 2 -c 2 -o gpurun_out/prof_${tag} -f \
    python tools/prof_step.py $wl $nm $K 3 > gpurun_out/ncu_f_${tag}.log 2>&1
tail -2 gpurun_out/ncu_f_${tag}.log
