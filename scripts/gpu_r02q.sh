#!/bin/bash
# final validation of the round-2 build: GPU tests, smoke, bench lines (both arms), the reference's benchmark programs,
# config 4 through b2WorldBatch
mkdir -p gpurun_out
TAG=r02q
( time python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in mixed_100k many_pyramids tumbler_worlds; do
  timeout 900 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$wl.json') if l.startswith('{')][0])
    r=d['roofline']
    print('$wl ms/step %.4f e2e %.4f frac %.3f frac_dram %s cpu ms/step %.2f'%(d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['frac_dram'], d['cpu_baseline']['ms_per_step']))
    print('   ', r['kernel_us_per_step'])
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/${TAG}_bench_$wl.err').read()[-1500:])
PY
done
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-260 gpurun_out/${TAG}_bench_reference.json
bash scripts/gpu_refbench.sh ${TAG} 2>&1 | tail -16
timeout 300 tests/cpp/build/batch_tumbler_gpu 256 500 100 > gpurun_out/${TAG}_batch_tumbler.txt 2>&1; tail -2 gpurun_out/${TAG}_batch_tumbler.txt
