#!/bin/bash
mkdir -p gpurun_out
GPSLC_BENCH_C5=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_n2c5.json 2> gpurun_out/bench_r02_n2c5.err
tail -4 gpurun_out/bench_r02_n2c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_n2c5.json')); print(d['value'], d['e2e']['value'], d['e2e']['gather_ms'], d['c5'], d['strong'])"
