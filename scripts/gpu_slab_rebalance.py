"""Slab decomposition WITH ownership migration on N GPUs (one process per GPU, torchrun): the sliding pile
of tests/test_slab.py::test_sliding_pile_migrates_between_slabs, halo exchange over NCCL every step, a
rebalance (all-gather of owner + contact records, re-cut, rebuild) every K steps.  Rank 0 prints the outcome
next to the single-arena run.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/gpu_slab_rebalance.py [bodies] [steps] [K]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from box2d_optimized_b200 import Arena, GpuScene, arena_from_scene
from box2d_optimized_b200.slab import (DistTransport, SlabRank, exchange_distributed, gather_records, make_slabs, rebalance_distributed,
                                       scene_arrays)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
K = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = GpuScene("mixed", n, 12345, device=local)
glob = scene_arrays(scene)
x_start = glob["bodies"][:, 4].copy()
slabs, owner0, cuts0 = make_slabs(glob, world, halo=3.0)
sr = SlabRank(glob, slabs[rank], device=local)
transport = DistTransport(rank, world, local)
P = Arena.params(gravity=(5.0, -10.0))
arrived, cuts = 0, cuts0
for k in range(steps):
    sr.arena.step(P, None)
    exchange_distributed(sr, transport)
    if (k + 1) % K == 0 and k + 1 < steps:
        sr, owner, cuts, a = rebalance_distributed(glob, sr, halo=3.0, device=local)
        arrived += a
recs = gather_records(sr.owned_record())
tot = torch.tensor([arrived], device="cuda")
dist.all_reduce(tot)
if rank == 0:
    single = arena_from_scene(scene, device=local)
    single.find_new_contacts()
    for k in range(steps):
        single.step(P, None)
    ref = single.download_bodies(what=("pos",))["pos"]
    dyn = glob["bodies"][:, 11] == 2
    xs = np.full(len(dyn), np.nan)
    ys = np.full(len(dyn), np.nan)
    owners = np.zeros(len(dyn), np.int32)
    for r in recs:
        g = r[:, 0].astype(np.int64)
        xs[g], ys[g] = r[:, 1], r[:, 2]
        owners[g] += 1
    slid = float(ref[dyn, 0].mean() - x_start[dyn].mean())
    out = dict(n_gpus=world, bodies=n, steps=steps, rebalance_every=K, ownership_changes=int(tot.item()),
               cut_first=[float(c) for c in cuts0], cut_last=[float(c) for c in cuts],
               every_body_has_one_owner=bool((owners[dyn] == 1).all()), slid_m=slid,
               mean_x_slabs=float(np.nanmean(xs[dyn])), mean_x_single=float(ref[dyn, 0].mean()),
               min_y=float(np.nanmin(ys[dyn])))
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/slab_rebalance.json", "w"))
dist.barrier()
dist.destroy_process_group()
