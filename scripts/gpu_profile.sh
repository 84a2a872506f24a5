#!/bin/bash
# ncu evidence: (1) launch list with device times for a short bench run, (2) one full capture of
# the dominant solver kernel and the broadphase traversal.  Numbers printed under ncu are not bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 6 --warmup 70 --profile-steps 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_solve_velocity$|k_bp_traverse|k_narrowphase|k_solve_position$" -s 400 -c 8 \
    -o gpurun_out/prof python bench.py --steps 4 --warmup 70 --profile-steps 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
