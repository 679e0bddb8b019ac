#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "many_feature or cov_build or chol_logpdf" 2>&1 | tail -8 > gpurun_out/r02_test6.log
cat gpurun_out/r02_test6.log
python tools/gpu_cov_build_bw.py 1024,12,256 1024,1,256 > gpurun_out/plain_cov6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cov_build_sym -s 3 -c 1 -o gpurun_out/prof_cov_d12_r02b \
    python tools/gpu_cov_build_bw.py 1024,12,256 > gpurun_out/ncu_cov12b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cov_build_sym -s 3 -c 1 -o gpurun_out/prof_cov_d1_r02b \
    python tools/gpu_cov_build_bw.py 1024,1,256 > gpurun_out/ncu_cov1b.log 2>&1
cat gpurun_out/plain_cov6.log
