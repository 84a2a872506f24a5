#!/bin/bash
timeout 900 python -m pytest tests/test_step_parity.py tests/test_batched_worlds.py -q -m gpu 2>&1 | tail -3
for k in 5 6 7; do B2G_KERNEL_TIMING=1 timeout 300 tests/cpp/build/bench_suite_gpu $k $k 2>&1 | cut -c1-120 | grep -E "BENCH|fused|big_" ; done
