#!/bin/bash
# final validation of the round-2 build
mkdir -p gpurun_out
TAG=${1:-final}
( time python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in mixed_100k many_pyramids tumbler_worlds; do
  timeout 900 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$wl.json') if l.startswith('{')][0])
    r=d['roofline']
    print('$wl ms/step %.4f p50 %.4f e2e %.4f frac %.3f cpu ms/step %.2f clocks %s'%(d['ms_per_step'], d['ms_per_step_p50'], d['e2e']['ms_per_step'], r['frac'], d['cpu_baseline']['ms_per_step'], d['clocks']))
    print('   ', r['kernel_us_per_step'])
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/${TAG}_bench_$wl.err').read()[-1500:])
PY
done
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
( time timeout 600 tests/cpp/build/bench_suite_gpu ) > gpurun_out/${TAG}_bench_suite_gpu.txt 2>&1
timeout 600 tests/cpp/build/single_gpu > gpurun_out/${TAG}_single_gpu.txt 2>&1; cat gpurun_out/${TAG}_single_gpu.txt
timeout 300 tests/cpp/build/batch_tumbler_gpu 1024 500 100 > gpurun_out/${TAG}_batch_tumbler_1024.txt 2>&1; tail -1 gpurun_out/${TAG}_batch_tumbler_1024.txt
