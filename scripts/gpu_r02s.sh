#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_api.py tests/test_batched_worlds.py -q -m gpu 2>&1 | tail -6
timeout 300 tests/cpp/build/batch_tumbler_gpu 256 500 100 2>&1 | tail -2
timeout 300 tests/cpp/build/batch_tumbler_gpu 1024 500 100 2>&1 | tail -2 | tee gpurun_out/r02s_batch_tumbler_1024.txt
