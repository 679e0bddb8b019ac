#!/bin/bash
# round-2 GPU batch 3: full tests, driver-style bench, cov-build bandwidth + write-only peak, c5 share, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_test3.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err ) 2> gpurun_out/bench_r02c.time
python tools/gpu_cov_build_bw.py > gpurun_out/cov_bw_r02c.log 2>&1
python tools/gpu_c5_sweep.py 32 8192 > gpurun_out/c5_share_r02c.json 2> gpurun_out/c5_share_r02c.err
# ncu: launch list of a short bench command, then full captures of the two kernels (the plain runs above exited first)
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/plain_short2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mh_lanes -s 3 -c 1 -o gpurun_out/prof_mh_r02 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/ncu_mh.log 2>&1
python tools/gpu_cov_build_bw.py 1024,12,256 1024,1,256 > gpurun_out/plain_cov.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cov_build_sym -s 3 -c 1 -o gpurun_out/prof_cov_d12_r02 \
    python tools/gpu_cov_build_bw.py 1024,12,256 > gpurun_out/ncu_cov12.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cov_build_sym -s 3 -c 1 -o gpurun_out/prof_cov_d1_r02 \
    python tools/gpu_cov_build_bw.py 1024,1,256 > gpurun_out/ncu_cov1.log 2>&1
tail -15 gpurun_out/r02_test3.log
cat gpurun_out/bench_r02c.time; cut -c1-600 gpurun_out/bench_r02c.json
cat gpurun_out/cov_bw_r02c.log; cat gpurun_out/c5_share_r02c.json; ls -la gpurun_out/*.ncu-rep
