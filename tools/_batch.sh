mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r02i.json 2> gpurun_out/bench_n8_r02i.err ) 2>&1 | tail -3
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8_r02i.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], d['e2e'].get('gather_ms'), 'strong', d.get('strong',{}).get('value'), 'c2', d['c2']['value'], 'c4', d['c4']['value'], 'c5', d.get('c5'))"
tail -3 gpurun_out/bench_n8_r02i.err
