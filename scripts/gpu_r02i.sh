#!/bin/bash
# full validation of the current build: GPU tests, smoke, the three bench workloads with CPU baseline, reference arm
mkdir -p gpurun_out
TAG=r02i
( time python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in mixed_100k many_pyramids tumbler_worlds; do
  timeout 900 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$wl.json') if l.startswith('{')][0])
    print('$wl ms/step %.4f e2e %.4f frac %.3f cpu %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('cpu_baseline')))
    print('   ', d['roofline']['kernel_us_per_step'])
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/${TAG}_bench_$wl.err').read()[-1500:])
PY
done
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-600 gpurun_out/${TAG}_bench_reference.json
