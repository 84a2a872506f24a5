#!/bin/bash
# One GPU round: tests, smoke, bench, launch list, one full ncu capture of the dominant kernel.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 200 --warmup 60 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 100 --warmup 60 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
