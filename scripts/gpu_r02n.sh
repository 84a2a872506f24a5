#!/bin/bash
mkdir -p gpurun_out
for sct in 6 4 3 2; do
  echo "== B2G_CUT_SECTORS=$sct"
  B2G_CUT_SECTORS=$sct B2G_CUDA_LIB=box2d_optimized_b200/libb2cuda_trace.so timeout 300 python scripts/gpu_big_trace.py 100000 310 2>&1 | grep -E "cut colour|velocity sweep [34] done|interior of sweep 4|position sweep 1|^  end|ms/step"
done
