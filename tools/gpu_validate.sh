#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_test25.log; cat gpurun_out/r02_test25.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err ) 2> gpurun_out/bench_r02n.time; tail -3 gpurun_out/bench_r02n.time
python -c "
import json; d=json.load(open('gpurun_out/bench_r02n.json')); print(round(d['value'],1), round(d['roofline']['frac'],4), 'c2', round(d['c2']['value']), round(d['c2']['frac'],3), 'c4', round(d['c4']['value'],2), round(d['c4']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'ite', round(d['ite']['value']), 'c1', d.get('c1'), d['cpu_baseline'].get('variants',{}).get('c1_seconds'), d['clocks'])"
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r02n.json 2> /dev/null ) 2>&1 | tail -3; cut -c1-200 gpurun_out/bench_ref_r02n.json
