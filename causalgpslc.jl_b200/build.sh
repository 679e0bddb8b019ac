#!/bin/bash
# Build libgpslc_b200.so (sm_100a only) in-tree. Usage: causalgpslc.jl_b200/build.sh
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/lib"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# GPSLC_EXTRA_FLAGS / GPSLC_LIB_SUFFIX: development builds (e.g. -DGPSLC_PHASE_TIMING -> libgpslc_b200_prof.so)
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 ${GPSLC_EXTRA_FLAGS:-}"
SUF="${GPSLC_LIB_SUFFIX:-}"
objs=""
pids=""
for src in "$here"/csrc/*.cu; do
  obj="$here/lib/$(basename "${src%.cu}")$SUF.o"
  stale=0
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then stale=1; fi
  for hdr in "$here"/csrc/*.cuh "$here"/../include/*.h; do
    if [ -f "$obj" ] && [ "$hdr" -nt "$obj" ]; then stale=1; fi
  done
  if [ $stale = 1 ]; then
    $NVCC $FLAGS -c "$src" -o "$obj" &
    pids="$pids $!"
  fi
  objs="$objs $obj"
done
for p in $pids; do wait $p; done
$NVCC -shared -o "$here/lib/libgpslc_b200$SUF.so" $objs -lcudart
echo "built $here/lib/libgpslc_b200$SUF.so"
