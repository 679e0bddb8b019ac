mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_r02i.json 2> gpurun_out/bench_n2_r02i.err ) 2>&1 | tail -3
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_r02i.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], d['e2e'].get('gather_ms'), 'strong', d.get('strong'), 'c2', d['c2']['value'], 'c1', d.get('c1',{}).get('seconds'))"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2>/dev/null ) 2>&1 | tail -3; cut -c1-160 gpurun_out/bench_ref_n2.json
tail -3 gpurun_out/bench_n2_r02i.err
