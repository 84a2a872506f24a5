#!/bin/bash
# sequenced cut constraints: targeted tests under a timeout (a wrong wait spins for ever), trace, suite, benches
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_step_parity.py -q -m gpu -x -k "tiled or colouring or determin" -s 2>&1 | tail -30 ) > gpurun_out/r02d_pytest_tiles.txt 2>&1; cat gpurun_out/r02d_pytest_tiles.txt
if grep -q "passed" gpurun_out/r02d_pytest_tiles.txt && ! grep -q "failed\|error" gpurun_out/r02d_pytest_tiles.txt; then
  B2G_CUDA_LIB=box2d_optimized_b200/libb2cuda_trace.so timeout 300 python scripts/gpu_big_trace.py 100000 310 > gpurun_out/r02d_tile_trace.txt 2>&1; head -30 gpurun_out/r02d_tile_trace.txt
  ( time timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/r02d_pytest_gpu.txt 2>&1; tail -15 gpurun_out/r02d_pytest_gpu.txt
  for wl in mixed_100k tumbler_worlds; do
    timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_$wl.json 2> gpurun_out/r02d_bench_$wl.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02d_bench_$wl.json') if l.startswith('{')][0])
    c=d['config']
    print('$wl ms/step %.4f p50 %.4f e2e %.4f launches/step %.1f colours %d serial %.0f rounds %.1f contacts %.0f'%(d['ms_per_step'],d['ms_per_step_p50'],d['e2e']['ms_per_step'],d['gpu_launches']/d['steps'],c['colours_max'],c['serial_bucket_constraints_mean'],c['colour_rounds_mean'],c['contacts_mean']))
    print('   ',d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['kernel_us_per_step'])
except Exception as e: print('$wl failed', e)
PY
    tail -3 gpurun_out/r02d_bench_$wl.err
  done
  B2G_TILE_BARRIERS=1 timeout 600 python bench.py --workload mixed_100k --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02d_bench_mixed_100k_barriers.json 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02d_bench_mixed_100k_barriers.json') if l.startswith('{')][0])
print('TILE BARRIERS mixed_100k ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_us_per_step'])
PY
fi
