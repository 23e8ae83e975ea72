#!/bin/bash
# usage: tools/gpu_cf.sh <tag>  -- closed-form tests + cfg2/cfg5 closed-form bench lines
tag=$1
python -m pytest tests/test_gpu_closed_form.py -x -q 2>&1 | tail -4
for wl in cfg2 cfg5; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl --numerics closed_form > gpurun_out/cf_${tag}_$wl.json 2> gpurun_out/cf_${tag}_$wl.err
  tail -1 gpurun_out/cf_${tag}_$wl.err
done
