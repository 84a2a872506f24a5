#!/bin/bash
# weak-scaling run at N GPUs (launched exactly as the driver does)
N=$1; WL=${2:-pyramid_worlds}; STEPS=${3:-100}; WARM=${4:-60}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --workload $WL --steps $STEPS --warmup $WARM --no-cpu-baseline 2> gpurun_out/scale_${WL}_$N.err | tail -1 | tee gpurun_out/scale_${WL}_$N.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $WL --steps $STEPS --warmup $WARM --no-cpu-baseline 2> gpurun_out/scale_${WL}_$N.err | tail -1 | tee gpurun_out/scale_${WL}_$N.json
fi
tail -3 gpurun_out/scale_${WL}_$N.err
