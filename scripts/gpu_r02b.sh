#!/bin/bash
# worklist colouring + world-local keys + host fixes: suite and benches
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 ) > gpurun_out/r02b_pytest_gpu.txt 2>&1; cat gpurun_out/r02b_pytest_gpu.txt
for wl in mixed_100k many_pyramids tumbler_worlds; do
  python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_$wl.json 2> gpurun_out/r02b_bench_$wl.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02b_bench_$wl.json') if l.startswith('{')][0])
print('$wl ms/step %.4f p50 %.4f e2e %.4f launches/step %.1f colours %d'%(d['ms_per_step'],d['ms_per_step_p50'],d['e2e']['ms_per_step'],d['gpu_launches']/d['steps'],d['config']['colours_max']))
print('   ',d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['kernel_us_per_step'])
PY
  tail -3 gpurun_out/r02b_bench_$wl.err
done
