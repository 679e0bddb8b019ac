#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_cov_build_bw.py > gpurun_out/cov_bw_r02d.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cov_build or summarize" 2>&1 | tail -5 > gpurun_out/r02_test4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
cat gpurun_out/cov_bw_r02d.log; cat gpurun_out/r02_test4.log; tail -5 gpurun_out/bench_r02_n2.err; cut -c1-300 gpurun_out/bench_r02_n2.json
