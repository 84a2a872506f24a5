#!/bin/bash
mkdir -p gpurun_out
for k in 12 5 4 7 10; do
  B2G_KERNEL_TIMING=1 timeout 300 tests/cpp/build/bench_suite_gpu $k $k ${1:-100} 2>&1 | cut -c1-200
done > gpurun_out/refbench_diag.txt
cat gpurun_out/refbench_diag.txt
