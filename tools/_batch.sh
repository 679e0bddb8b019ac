mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_eg.json 2> gpurun_out/bench_eg.err
python -c "
import json; d=json.load(open('gpurun_out/bench_eg.json')); print(round(d['value'],1), round(d['roofline']['frac'],4), 'c2', round(d['c2']['value']), round(d['c2']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'c1', d.get('c1',{}).get('seconds'), d.get('c1',{}).get('gpslc_seconds'))"
for g in 1 2 4 8; do echo group $g; GPSLC_ESS_GROUP=$g python tools/gpu_ess_pass_time.py 256 4 5 1024 2>&1 | tail -1; GPSLC_ESS_GROUP=$g python tools/gpu_ess_pass_time.py 150 6 0 1 2>&1 | tail -1; GPSLC_ESS_GROUP=$g python tools/gpu_ess_pass_time.py 1024 16 10 512 2>&1 | tail -1; done
