#!/bin/bash
# usage: tools/bench_variants.sh <tag> [bench args]  -- one bench line per library in variants/
tag=$1; shift
for so in variants/*.so; do
  n=$(basename $so .so)
  SRM_PHYSICS_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bv_${tag}_$n.json 2> gpurun_out/bv_${tag}_$n.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bv_${tag}_$n.json"))
    print("$n", "ms/step %.4f fwd %.4f bwd %.4f frac %.4f" % (d["ms_per_step"], d["roofline"]["fwd_ms"], d["roofline"]["bwd_ms"], d["roofline"]["frac"]))
except Exception as e:
    print("$n", "FAILED", e)
PY
done
