#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_host_api.py tests/test_config_sizes.py tests/test_reference_benchmarks.py tests/test_batched_worlds.py tests/test_step_parity.py -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r02m_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r02m_pytest_gpu.txt
python scripts/gpu_tumbler_stats.py | tail -1
