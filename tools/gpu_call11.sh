#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_test11.log; cat gpurun_out/r02_test11.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r02e_rcp.json 2> gpurun_out/bench_r02e_rcp.err
GPSLC_LIB_SUFFIX=_norcp python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r02e_norcp.json 2> gpurun_out/bench_r02e_norcp.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r02e_rcp2.json 2> /dev/null
for f in rcp norcp rcp2; do python -c "
import json; d=json.load(open('gpurun_out/bench_r02e_$f.json')); print('$f', round(d['value'],1), round(d['roofline']['frac'],4), round(d['c2']['value']), round(d['c2']['frac'],3), round(d['c4']['value'],2), round(d['c4']['frac'],3))"; done
