for shape in "256 4 5 1024" "1024 16 10 512" "128 4 4 2048"; do
  for suf in "" _ld68 _ld72 ""; do GPSLC_LIB_SUFFIX=$suf python tools/gpu_small_n_time.py $shape 2>&1 | grep sweeps; done
done
