#!/bin/bash
# ncu evidence: (1) launch list with device times for a short bench run, (2) one full capture of
# the dominant kernels.  Numbers printed under ncu are not bench values.
# -s counts ALL launches, not only the ones the -k filter keeps: at 15 launches per step (many_pyramids) 1350 is
# step 90, inside the settled window; with -s 400 the r01g capture landed on a step with at most 2 048 contact slots
# (profiles/r01g_ncu_full_early_step.csv).
mkdir -p gpurun_out
WL=${1:-many_pyramids}
SKIP=${2:-900}
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --workload $WL --steps 8 --warmup 70 --profile-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_solve_bins_fused|k_bp_traverse|k_narrowphase|k_wide_refit|k_island_union$" -s ${3:-1350} -c 10 \
    -o gpurun_out/prof python bench.py --workload $WL --steps 4 --warmup 100 --profile-steps 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_bench.log | cut -c1-200
ls -la gpurun_out/
