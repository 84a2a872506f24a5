#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_step_parity.py -q -m gpu -x -k "tiled or determin" > gpurun_out/r02e_pytest_tiles.txt 2>&1; rc=$?; tail -3 gpurun_out/r02e_pytest_tiles.txt
if [ $rc -eq 0 ]; then
  B2G_CUDA_LIB=box2d_optimized_b200/libb2cuda_trace.so timeout 300 python scripts/gpu_big_trace.py 100000 310 2>&1 | head -24 > gpurun_out/r02e_tile_marks.txt; cat gpurun_out/r02e_tile_marks.txt
  for wl in mixed_100k; do
    timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_$wl.json 2> gpurun_out/r02e_bench_$wl.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02e_bench_$wl.json') if l.startswith('{')][0])
    c=d['config']
    print('$wl ms/step %.4f p50 %.4f e2e %.4f launches/step %.1f colours %d serial %.0f rounds %.1f contacts %.0f'%(d['ms_per_step'],d['ms_per_step_p50'],d['e2e']['ms_per_step'],d['gpu_launches']/d['steps'],c['colours_max'],c['serial_bucket_constraints_mean'],c['colour_rounds_mean'],c['contacts_mean']))
    print('   ',d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['kernel_us_per_step'])
except Exception as e: print('$wl failed', e)
PY
    tail -3 gpurun_out/r02e_bench_$wl.err
  done
fi
