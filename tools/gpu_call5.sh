#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ite or sate or counterfactual or summarize or public_api" 2>&1 | tail -30 > gpurun_out/r02_test5a.log
tail -5 gpurun_out/r02_test5a.log
timeout 300 python tools/gpu_c5_sweep.py 32 8192 > gpurun_out/c5_share_on.json 2> gpurun_out/c5_share_on.err
GPSLC_ITE_SHARE=0 timeout 300 python tools/gpu_c5_sweep.py 32 8192 > gpurun_out/c5_share_off.json 2> gpurun_out/c5_share_off.err
cat gpurun_out/c5_share_on.json gpurun_out/c5_share_off.json; tail -3 gpurun_out/c5_share_on.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_test5b.log
tail -6 gpurun_out/r02_test5b.log
