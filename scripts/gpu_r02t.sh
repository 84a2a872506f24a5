#!/bin/bash
timeout 900 python -m pytest tests/test_host_api.py tests/test_batched_worlds.py tests/test_reference_benchmarks.py tests/test_joint_parity.py -q -m gpu 2>&1 | tail -15
tests/cpp/build/api_gpu | grep churn
