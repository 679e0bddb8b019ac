#!/bin/bash
# round-2 GPU batch: tests, fold A/B, cov-build bandwidth, c5 team/CTA knobs (one GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_test2.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r02b_fold.json 2> gpurun_out/bench_r02b_fold.err
GPSLC_LIB_SUFFIX=_nofold python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r02b_nofold.json 2> gpurun_out/bench_r02b_nofold.err
python tools/gpu_cov_build_bw.py > gpurun_out/cov_bw_r02b.log 2>&1
python tools/gpu_c5_sweep.py 32 8192 > gpurun_out/c5_default.json 2> gpurun_out/c5_default.err
GPSLC_TEAM=4 GPSLC_CTAS_PER_SM=1 python tools/gpu_c5_sweep.py 32 8192 > gpurun_out/c5_t4c1.json 2> gpurun_out/c5_t4c1.err
GPSLC_TEAM=8 GPSLC_CTAS_PER_SM=1 python tools/gpu_c5_sweep.py 32 8192 > gpurun_out/c5_t8c1.json 2> gpurun_out/c5_t8c1.err
GPSLC_TEAM=4 GPSLC_CTAS_PER_SM=2 python tools/gpu_c5_sweep.py 32 8192 > gpurun_out/c5_t4c2.json 2> gpurun_out/c5_t4c2.err
cat gpurun_out/r02_test2.log | tail -30
cat gpurun_out/bench_r02b_fold.json gpurun_out/bench_r02b_nofold.json | cut -c1-400
cat gpurun_out/cov_bw_r02b.log
cat gpurun_out/c5_*.json
