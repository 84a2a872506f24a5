#!/bin/bash
one() {
  env $1 timeout 300 python bench.py --workload $2 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1 $2 %.4f p50 %.4f'%(d['ms_per_step'],d['ms_per_step_p50']), d['roofline']['kernel_us_per_step'])"
}
one B2G_CUDA_LIB=box2d_optimized_b200/libb2cuda_t4.so mixed_100k
