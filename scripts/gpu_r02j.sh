#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_joint_parity.py tests/test_host_api.py tests/test_batched_worlds.py tests/test_reference_benchmarks.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r02j_pytest_joints.txt; tail -30 gpurun_out/r02j_pytest_joints.txt
