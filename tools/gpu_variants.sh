#!/bin/bash
# timing experiments: per-layer times of one forward with alternative builds of the library
mkdir -p gpurun_out
for f in gpurun_variants/libpds_*.so; do
  v=$(basename $f .so); v=${v#libpds_}
  PDS_B200_LIB=$PWD/$f PDS_B200_PROFILE_DETAIL=1 timeout 300 python tools/bench_detail.py 2>&1 | head -${1:-4} > gpurun_out/variant_$v.txt
  echo "== $v"; cat gpurun_out/variant_$v.txt
done
