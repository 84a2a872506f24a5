#!/bin/bash
timeout 300 python bench.py --workload tumbler_worlds --steps 30 --warmup 5 --no-cpu-baseline --no-e2e | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('tumbler_worlds', d['ms_per_step'], d['roofline']['kernel_us_per_step'])"
timeout 900 python -m pytest tests/test_step_parity.py tests/test_batched_worlds.py tests/test_joint_parity.py tests/test_config_sizes.py -q -m gpu 2>&1 | tail -4
for k in 5 6 7; do B2G_KERNEL_TIMING=1 timeout 300 tests/cpp/build/bench_suite_gpu $k $k 2>&1 | cut -c1-120 | grep -E "BENCH|fused|big_" ; done
