#!/bin/bash
for v in xGEN xDIAG xEPI ""; do
  GPSLC_LIB_SUFFIX=${v:+_$v} timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_exp_$v.json 2> gpurun_out/bench_exp_$v.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/bench_exp_$v.json')); print('variant [$v]', round(d['value'],1), round(d['c2']['value']), round(d['c4']['value'],2))
except Exception as e: print('variant [$v] failed', e)"
  tail -2 gpurun_out/bench_exp_$v.err
done
