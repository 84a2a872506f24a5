#!/bin/bash
# full GPU suite + the three headline benches (short form)
TAG=${1:-r02}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -25 gpurun_out/${TAG}_pytest_gpu.txt
for wl in mixed_100k many_pyramids tumbler_worlds; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$wl.json') if l.startswith('{')][0])
    c=d['config']
    print('$wl ms/step %.4f p50 %.4f e2e %.4f launches/step %.1f colours %d serial %.0f rounds %.1f contacts %.0f'%(d['ms_per_step'],d['ms_per_step_p50'],d['e2e']['ms_per_step'],d['gpu_launches']/d['steps'],c['colours_max'],c['serial_bucket_constraints_mean'],c['colour_rounds_mean'],c['contacts_mean']))
    print('   ',d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['kernel_us_per_step'])
except Exception as e: print('$wl failed', e)
PY
  tail -3 gpurun_out/${TAG}_bench_$wl.err
done
