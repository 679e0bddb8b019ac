#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_test20.log; cat gpurun_out/r02_test20.log
python tools/gpu_small_n_time.py 256 4 5 1024 2>&1 | tail -2 | head -1; python tools/gpu_small_n_time.py 150 6 0 2048 2>&1 | tail -2 | head -1
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02f.json')); print(round(d['value'],1), round(d['roofline']['frac'],4), 'c2', round(d['c2']['value']), round(d['c2']['frac'],3), 'c4', round(d['c4']['value'],2), 'e2e', round(d['e2e']['value'],1), 'ite', round(d['ite']['value']))"
