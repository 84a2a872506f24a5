#!/bin/bash
# one full ncu capture (source-level stall reasons) of the oversize-island kernel on the settled 100k-body pile
mkdir -p gpurun_out
WL=${1:-mixed_100k}
KERN=${2:-k_big_tiles}
SKIP=${3:-302}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s $SKIP -c 1 -o gpurun_out/ncu_$KERN -f \
    python bench.py --workload $WL --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --no-roofline > gpurun_out/ncu_$KERN.log 2>&1
tail -3 gpurun_out/ncu_$KERN.log | cut -c1-300
ls -la gpurun_out/ncu_$KERN.ncu-rep
