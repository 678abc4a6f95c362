#!/bin/bash
# ncu --set full capture of the block-ordered kernels (k_bx) on the C4 shard
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r2_bx}
for f in 1 0; do
  TB_FUSE_LHS=$f timeout 600 ncu --set full --clock-control none --import-source on \
    --kernel-name-base demangled -k "regex:k_bx<" -c 3 -f -o $OUT/prof_${TAG}_f$f python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/prof_${TAG}_f$f.log 2>&1
  tail -3 $OUT/prof_${TAG}_f$f.log
done
ls -la $OUT/*.ncu-rep
