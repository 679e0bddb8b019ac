for shape in "1024 16 10 512" "2048 32 10 128"; do
  for suf in "" _pf6 _pf12 _pf24 ""; do GPSLC_LIB_SUFFIX=$suf python tools/gpu_small_n_time.py $shape 2>&1 | grep sweeps; done
done
