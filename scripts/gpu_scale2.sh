#!/bin/bash
# multi-GPU runs (call with gpurun --gpus N): slab decomposition (config 5) and batched tumbler worlds (config 4)
N=${1:-2}
TAG=${2:-r02h}
mkdir -p gpurun_out
run() {  # workload, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --workload $1 --steps 30 --warmup 5 --no-cpu-baseline $2 > gpurun_out/${TAG}_scale_$1_$N.json 2> gpurun_out/${TAG}_scale_$1_$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_scale_$1_$N.json') if l.startswith('{')][0])
    print('$1 N=$N ms/step %.4f value %.1fM e2e %s halo %s contacts %.0f'%(d['ms_per_step'], d['value']/1e6, d['e2e'] and round(d['e2e']['ms_per_step'],4), d['config']['halo_bytes_per_step_per_rank'], d['config']['contacts_mean']))
except Exception as e:
    print('$1 N=$N failed', e); print(open('gpurun_out/${TAG}_scale_$1_$N.err').read()[-1500:])
PY
}
for wl in "$@"; do :; done
shift 2
for wl in "$@"; do run $wl ""; done
