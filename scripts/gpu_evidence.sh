#!/bin/bash
# ncu counter evidence for the final build: per-kernel sections (speed of light, memory workload,
# occupancy, warp-state stall breakdown, scheduler) in the settled windows of mixed_100k and many_pyramids, launch
# lists of the same commands, and compute-sanitizer memcheck + racecheck logs.  bench.py --profiler-range brackets
# the timed steps with cudaProfilerStart/Stop, so every capture is of the timed window.  Numbers printed under ncu
# are not bench values.
mkdir -p gpurun_out
TAG=${1:-r02j}
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section SchedulerStats --section LaunchStats --section ComputeWorkloadAnalysis"
B="--no-cpu-baseline --no-e2e --no-roofline --profiler-range"
KERN="k_bp_traverse|k_narrowphase|k_colour_worklist|k_island_union|k_island_flatten|k_mark_active_bins|k_solve_bins_fused|k_tile_assign|k_wide_refit|k_island_alloc|k_contact_sweep|k_contact_insert"
for wl in mixed_100k many_pyramids; do
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_$wl.csv \
      python bench.py --workload $wl --steps 20 --warmup 5 $B > gpurun_out/${TAG}_ncu_launches_$wl.log 2>&1
  timeout 900 ncu $SECS --clock-control none --profile-from-start off -k regex:"$KERN" -s 24 -c 12 \
      -o gpurun_out/${TAG}_$wl -f python bench.py --workload $wl --steps 4 --warmup 5 $B > gpurun_out/${TAG}_ncu_$wl.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_$wl.log | cut -c1-200
done
timeout 600 ncu $SECS --clock-control none --profile-from-start off -k regex:"k_big_tiles" -s 2 -c 1 \
    -o gpurun_out/${TAG}_big_tiles -f python bench.py --workload mixed_100k --steps 4 --warmup 5 $B > gpurun_out/${TAG}_ncu_big_tiles.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_big_tiles.log | cut -c1-200
if [ ! -s gpurun_out/${TAG}_big_tiles.ncu-rep ]; then
  # kernel replay could not restore the state of the grid-barrier kernel: collect the counters by re-running the
  # (deterministic) application once per pass instead
  timeout 1200 ncu $SECS --replay-mode application --clock-control none --profile-from-start off -k regex:"k_big_tiles" -s 2 -c 1 \
      -o gpurun_out/${TAG}_big_tiles -f python bench.py --workload mixed_100k --steps 4 --warmup 5 $B > gpurun_out/${TAG}_ncu_big_tiles_app.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_big_tiles_app.log | cut -c1-200
fi
( time timeout 900 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py 1.0 ) > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.txt
( time timeout 900 compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py 0.25 ) > gpurun_out/${TAG}_sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.txt
ls -la gpurun_out/*.ncu-rep
