#!/bin/bash
# GC checks + cfg4-shaped timing (K realisations per GPU given as $1, default 2)
K=${1:-2}
python -m pytest tests/test_gpu_gc.py -m gpu -q 2>&1 | tail -5
for lut in "" "--no-pvt-lut"; do
  python bench.py --workload cfg4 --K $K --steps 5 --warmup 3 --no-cpu-baseline $lut 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg4 K=$K lut=\"$lut\"', 'ms', d['ms_per_step'], 'fwd', d['roofline']['fwd_ms'], 'bwd', d['roofline']['bwd_ms'], 'value %.3e'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'])
    else: print(l.rstrip()[:300])
"
done
