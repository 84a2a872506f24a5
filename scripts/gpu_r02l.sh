#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r02l_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r02l_pytest_gpu.txt
( time timeout 400 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py 0.25 ) > gpurun_out/r02l_sanitizer_memcheck.txt 2>&1; tail -6 gpurun_out/r02l_sanitizer_memcheck.txt
( time timeout 200 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py 0.75 mixed ) > gpurun_out/r02l_sanitizer_memcheck_bigpath.txt 2>&1; tail -6 gpurun_out/r02l_sanitizer_memcheck_bigpath.txt
( time timeout 300 compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py 0.25 ) > gpurun_out/r02l_sanitizer_racecheck.txt 2>&1; tail -6 gpurun_out/r02l_sanitizer_racecheck.txt
