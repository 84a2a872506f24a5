#!/bin/bash
mkdir -p gpurun_out
WL=$1; WARM=$2; SKIP=$3
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 600 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --workload $WL --steps 6 --warmup $WARM --profile-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench_$WL.log 2>&1
tail -2 gpurun_out/ncu_bench_$WL.log | cut -c1-300
