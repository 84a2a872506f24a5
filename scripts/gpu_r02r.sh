#!/bin/bash
B2G_CUDA_LIB=box2d_optimized_b200/libb2cuda_trace.so timeout 300 python scripts/gpu_big_trace.py 100000 310 --tiles 2>&1 | grep -E -A19 "tile interior time" | grep -v "^barriers\|^per barrier"
timeout 300 python bench.py --workload mixed_100k --steps 40 --warmup 5 --no-cpu-baseline --no-e2e | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('mixed_100k %.4f p50 %.4f'%(d['ms_per_step'], d['ms_per_step_p50']), 'frac %.3f'%d['roofline']['frac'], d['roofline']['kernel_us_per_step'])"
timeout 600 python -m pytest tests/test_step_parity.py -q -m gpu -x -k "tile or determin" 2>&1 | tail -2
