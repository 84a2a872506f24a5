#!/bin/bash
# the reference's own benchmark programs on the GPU path and on the box's CPU, side by side
mkdir -p gpurun_out
TAG=${1:-r02i}
( time timeout 900 tests/cpp/build/bench_suite_gpu ) > gpurun_out/${TAG}_bench_suite_gpu.txt 2>&1
( time timeout 900 tests/cpp/build/bench_suite_ref ) > gpurun_out/${TAG}_bench_suite_ref.txt 2>&1
timeout 900 tests/cpp/build/single_gpu > gpurun_out/${TAG}_single_gpu.txt 2>&1
timeout 900 tests/cpp/build/single_ref > gpurun_out/${TAG}_single_ref.txt 2>&1
python - <<PY
import sys
sys.path.insert(0,'tests')
from test_reference_benchmarks import parse
g=parse(open('gpurun_out/${TAG}_bench_suite_gpu.txt').read()); r=parse(open('gpurun_out/${TAG}_bench_suite_ref.txt').read())
for k in sorted(r):
    if k not in g: print('b%d missing on gpu'%k); continue
    size=max(1.0,r[k]['q90_y']-r[k]['q10_y'])
    print('b%-2d %-30s cpu %8.0f ms gpu %8.0f ms x%5.1f | awake %d/%d contacts %d/%d | d/size mean_y %.3f med %.3f q10 %.3f q90 %.3f mean_x %.3f'%(
        k,r[k]['name'],r[k]['total_ms'],g[k]['total_ms'],r[k]['total_ms']/g[k]['total_ms'],g[k]['awake'],r[k]['awake'],g[k]['contacts'],r[k]['contacts'],
        *[abs(g[k][q]-r[k][q])/size for q in ('mean_y','med_y','q10_y','q90_y','mean_x')]))
PY
tail -3 gpurun_out/${TAG}_bench_suite_gpu.txt | cut -c1-300
