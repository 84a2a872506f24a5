#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_narrowphase_parity.py tests/test_world_step_parity.py tests/test_host_api.py tests/test_step_parity.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02k_pytest.txt; tail -15 gpurun_out/r02k_pytest.txt
timeout 300 tests/cpp/build/batch_tumbler_gpu 256 500 100 2>&1 | tail -3
one() {
  env $1 timeout 300 python bench.py --workload mixed_100k --steps 40 --warmup 5 --no-cpu-baseline --no-e2e | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'mixed_100k %.4f'%d['ms_per_step'], 'rounds', d['config'].get('colour_rounds_mean'), d['roofline']['kernel_us_per_step'])"
}
one B2G_X=0
one B2G_WL_SINGLE_MAX=8192
one B2G_WL_SINGLE_MAX=32768
