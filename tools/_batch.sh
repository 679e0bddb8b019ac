mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print(round(d['value'],1), round(d['roofline']['frac'],4), 'c2', round(d['c2']['value']), round(d['c2']['frac'],3), 'c4', round(d['c4']['value'],2), round(d['c4']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'ite', round(d['ite']['value']), 'c1', d.get('c1',{}).get('seconds'))"
timeout 800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
