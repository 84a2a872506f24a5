#!/bin/bash
# round-2 first GPU call: suite, smoke, the new bench defaults on both arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
nproc >> gpurun_out/r02a_smi.txt
( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/r02a_pytest_gpu.txt 2>&1; cat gpurun_out/r02a_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02a_bench_mixed100k.json 2> gpurun_out/r02a_bench.err; tail -c 3000 gpurun_out/r02a_bench_mixed100k.json; tail -5 gpurun_out/r02a_bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02a_bench_ref_mixed100k.json 2>> gpurun_out/r02a_bench.err; cat gpurun_out/r02a_bench_ref_mixed100k.json; tail -4 gpurun_out/r02a_bench.err
( time python bench.py --workload many_pyramids --steps 100 --warmup 10 --no-cpu-baseline ) > gpurun_out/r02a_bench_many_pyramids.json 2>> gpurun_out/r02a_bench.err; tail -c 2500 gpurun_out/r02a_bench_many_pyramids.json
( time python bench.py --workload tumbler_worlds --steps 20 --warmup 5 ) > gpurun_out/r02a_bench_tumbler.json 2>> gpurun_out/r02a_bench.err; tail -c 2500 gpurun_out/r02a_bench_tumbler.json; tail -4 gpurun_out/r02a_bench.err
