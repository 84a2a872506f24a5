#!/usr/bin/env python
"""bench.py — ms per b2World::Step and body-steps/s (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload NAME]

A "step" is one b2World::Step of the workload.  At N=1 the workload is BASELINE.json configs[1]
("many_pyramids": 100 independent 20-row pyramids, 21 001 bodies, one world); for N>1 every rank
steps its own copy of that world (independent worlds sharded across GPUs, no data-path
collective: weak scaling) and `value` is the sum over ranks divided by the max-over-ranks time.

  value  : device-resident throughput — bodies x K / sum of per-step CUDA-event times on the
           arena's stream; L2 is flushed (256 MiB memset) between timed steps.
  e2e    : the same metric through the C-ABI with HOST buffers: every step uploads the force
           accumulators from pinned host memory (b2g_upload_forces), steps and reads every body's
           transform + velocity back to pinned host memory (b2g_step_download), all inside the
           timed region, one step at a time (no pipelining across steps).
  roofline     : the dominant KERNEL (largest CUDA-event time per launch over a profiled pass; the
                 library times every launch by kernel class), algorithmic bytes from SURVEY.md
                 §8(d) (table in DESIGN.md) / measured launch time, against MEASURED_PEAKS.json's
                 HBM copy bandwidth; `traffic` = DRAM bytes per launch of that kernel from the newest
                 committed ncu --set full summary under profiles/.
  cpu_baseline : the reference's own CPU Step (oracle/_ref, compiled from /root/reference) on a
                 bounded sample of the same workload, rank 0 / N=1 only.
`--impl reference` times that CPU implementation alone (one world per thread, n_gpus worlds).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic bytes per unit of work, SURVEY.md §8(d) (restated in DESIGN.md §5)
ALGO_BYTES = {
    "narrowphase": 324.0,       # per contact, box-box
    "islands": 28.0,            # per body / contact visited
    "integrate": 62.0,          # per body (60 velocity, 64 position pass)
    "colour": 16.0,             # per constraint per round (not in the survey: key + 2 body masks)
    "prepare": 292.0,           # per constraint
    "warm_start": 180.0,        # per constraint
    "solve_velocity": 196.0,    # per constraint per iteration (2-point manifold)
    "solve_position": 136.0,    # per constraint per iteration
    "store_impulses": 32.0,     # per constraint
    "finalize": 92.0,           # per body (64 write-back + 28 sleep)
    "bp_build": 144.0,          # per fixture (AABB + key + node build), spread over 5 kernels
    "bp_traverse": 32.0 * 15,   # per fixture: 32 B x log2(N_f) traversal upper bound
    "contact_merge": 32.0,      # per pair
    "sort_scan": 64.0,          # per key (4-pass radix of 8-byte key/value)
    # whole island solve per constraint per step at 8/3 iterations: prepare 292 + warm start 180 +
    # 8 x 196 + 3 x 136 + store 32 (the N_t term of SURVEY §8d's B_step)
    "fused_solve": 2480.0,
}

WORKLOADS = {
    # name: (scene, size, seed, description)
    "many_pyramids": ("many_pyramids", 100, 0, "100 independent 20-row pyramids (21001 bodies) in one world"),
    "pyramid": ("pyramid", 20, 0, "testbed pyramid, 20 rows (211 bodies)"),
    "mixed_100k": ("mixed", 100000, 12345, "100k circles + convex polygons settling into a container, sleeping on"),
    "mixed_10k": ("mixed", 10000, 12345, "10k circles + convex polygons settling into a container"),
    # batched independent worlds in ONE arena per GPU (world id in the broadphase key), sharded
    # round-robin over ranks: size = worlds per GPU
    "pyramid_worlds": ("pyramid", 20, 0, "512 independent 20-row pyramid worlds per GPU (108k bodies), one arena"),
    # BASELINE config 4 shape: motor-driven tumbler (benchmarks.h b3) with 500 boxes per world; the
    # scene is stepped to the state where every box has been spawned, then replicated per world
    "tumbler_worlds": ("tumbler", 500, 0, "config 4: 1024 independent 500-box tumbler worlds per GPU (8192 on 8 GPUs; 514k bodies per arena)"),
    # BASELINE config 5 shape: ONE world cut into x-slabs, one per GPU, ghost layer refreshed by a
    # per-step NCCL halo exchange (strong scaling: the world is fixed, the slab shrinks with N)
    "mixed_slab_400k": ("mixed", 400000, 12345, "one 400k-body world, x-slab per GPU, per-step NCCL halo exchange"),
    "mixed_slab_1m": ("mixed", 1000000, 12345, "one 1M-body world, x-slab per GPU, per-step NCCL halo exchange"),
}
SLAB = {"mixed_slab_400k", "mixed_slab_1m"}
WORLDS_PER_GPU = {"pyramid_worlds": 512, "tumbler_worlds": 1024}
PRESTEP = {"tumbler_worlds": 520}  # steps run through the drop-in API before the state is replicated


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="many_pyramids", choices=sorted(WORKLOADS))
    ap.add_argument("--worlds-per-gpu", type=int, default=0,
                    help="batched-world workloads: independent worlds per GPU (config 4 = 8192 worlds / 8 GPUs = 1024)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-steps", type=int, default=150)
    ap.add_argument("--profile-steps", type=int, default=20)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clocks and throttle reasons with nvidia-smi while the timed region runs"""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.stop_flag = threading.Event()
        self.samples = []
        self.reasons = set()
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU b2World::Step, one world per host thread."""
    if rank != 0:
        return
    from oracle.bindings import RefScene
    scene, size, seed, desc = WORKLOADS[args.workload]
    # one world per GPU of our arm for the single-world workloads; for the batched workloads a
    # bounded sample of one world per host thread (the per-world cost is identical)
    nworlds = max(1, args.gpus)
    ncpu = os.cpu_count() or 1
    if args.workload in WORLDS_PER_GPU:
        nworlds = min(ncpu, WORLDS_PER_GPU[args.workload] * max(1, args.gpus))
    nworlds = min(nworlds, max(1, ncpu))
    worlds = [RefScene(scene, size, seed) for _ in range(nworlds)]
    if args.workload in PRESTEP:
        [w.step(PRESTEP[args.workload]) for w in worlds]
    nb = worlds[0].body_count
    times = [0.0] * nworlds

    def work(i):
        worlds[i].step(args.warmup)
        times[i] = worlds[i].time_steps(args.steps)

    th = [threading.Thread(target=work, args=(i,)) for i in range(nworlds)]
    [t.start() for t in th]
    [t.join() for t in th]
    ms = max(times)
    value = nb * nworlds * args.steps / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "bodies_per_world": nb, "worlds": nworlds,
                   "velocity_iterations": 8, "position_iterations": 3, "dt": 1.0 / 60.0, "sleeping": True,
                   "continuous": False},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": nworlds, "kind": "reference",
                         "sample": f"steps {args.warmup}..{args.warmup + args.steps} of {args.workload}, "
                                   f"{nworlds} world(s), one thread each (the reference is single-threaded)"},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


KERNEL_OF_CLASS = {"fused_solve": "k_solve_bins_fused", "bp_traverse": "k_bp_traverse", "narrowphase": "k_narrowphase",
                   "solve_velocity": "k_big_solve"}


def ncu_traffic(cls, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest
    committed `ncu --set full` summary (profiles/*_ncu_full_summary.csv, written by
    scripts/summarize_ncu.py from a capture of the default workload); None when there is none."""
    import csv
    import glob
    if workload != "many_pyramids" or cls not in KERNEL_OF_CLASS:
        return None, None
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_ncu_full_summary.csv")))
    if not files:
        return None, None
    # newest summary that actually holds the dominant kernel (a truncated or foreign file must not
    # take the bench line down)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for f in reversed(files):
        try:
            rows = list(csv.reader(open(f)))
            hdr, units = rows[0], rows[1]
            ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            vals = [float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
                    for r in rows[2:] if len(r) > wi and KERNEL_OF_CLASS[cls] in r[0]]
        except (IndexError, ValueError, KeyError):
            continue
        if vals:
            return sum(vals) / len(vals), os.path.basename(f)
    return None, None


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from box2d_optimized_b200 import capi, Arena, arena_from_scene, GpuScene

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene_name, size, seed, desc = WORKLOADS[args.workload]
    lib = capi.load_cuda()
    # host-side scene construction through the drop-in C++ API; never stepped itself
    scene = GpuScene(scene_name, size, seed)
    if args.workload in PRESTEP:
        scene.step(PRESTEP[args.workload])
    nb = scene.body_count
    cap_contacts = max(4096, 8 * nb)

    from box2d_optimized_b200.sharding import aggregate, shard_worlds
    per_gpu = WORLDS_PER_GPU.get(args.workload, 1)
    if args.worlds_per_gpu > 0 and args.workload in WORLDS_PER_GPU:
        per_gpu = args.worlds_per_gpu
    my_worlds = shard_worlds(per_gpu * world, rank, world)   # independent worlds, no data-path collective
    copies = len(my_worlds)
    nb_world = nb
    nb = nb_world * copies
    cap_contacts = max(4096, 8 * nb)

    slab_mode = args.workload in SLAB and world > 1
    halo_bytes = 0
    if slab_mode:
        from box2d_optimized_b200.slab import SlabRank, exchange_distributed, make_slabs, scene_arrays
        glob = scene_arrays(scene)
        slabs, _, _ = make_slabs(glob, world, halo=3.0)
        nb_total = nb_world
        nb = slabs[rank].num_owned  # bodies this rank advances (ghosts are redundant work)
        copies = 1

    def fresh_arena():
        if slab_mode:
            sr = SlabRank(glob, slabs[rank], device=local_rank, max_contacts=max(4096, 8 * len(slabs[rank].global_ids)))
            sr.arena._slab = sr
            return sr.arena
        A = arena_from_scene(scene, max_contacts=cap_contacts, device=local_rank, copies=copies,
                             num_worlds=copies)
        A.find_new_contacts()
        return A

    def after_step(A):
        # the once-per-step halo exchange of the slab decomposition (NCCL point-to-point)
        if slab_mode:
            A.synchronize()
            exchange_distributed(A._slab)

    P = Arena.params()
    st = capi.StepStats()

    # ------------------------------------------------------------------ device-resident value
    A = fresh_arena()
    ext = torch.cuda.ExternalStream(A.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    for _ in range(args.warmup):
        A.step(P, st)
        after_step(A)
    A.synchronize()
    if slab_mode:
        halo_bytes = A._slab.halo_bytes()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches = 0
    contacts_seen, constraints_seen, colours_seen = [], [], []
    for k in range(args.steps):
        with torch.cuda.stream(ext):
            flush.zero_()  # evict the step's working set from L2 (outside the timed bracket)
        starts[k].record(ext)
        A.step(P, st)
        if slab_mode:
            after_step(A)
            ends[k].record()      # the exchange runs on torch's stream, after the arena's stream drained
        else:
            ends[k].record(ext)
        launches += st.num_launches
        contacts_seen.append(st.num_contacts)
        constraints_seen.append(st.num_constraints)
        colours_seen.append(st.num_colours)
    A.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.result()
    per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    elapsed_ms = sum(per_step_ms)
    _, elapsed_ms_max, value = aggregate(nb * args.steps, elapsed_ms, device=f"cuda:{local_rank}")

    # ------------------------------------------------------------------ per-kernel roofline pass
    A.set_kernel_timing(True)
    for _ in range(args.profile_steps):
        A.step(P, st)
        after_step(A)
    kt = A.kernel_timing()
    A.set_kernel_timing(False)
    total_kernel_ms = sum(v[0] for v in kt.values()) or 1.0
    # the dominant KERNEL: classes such as "colour" or "islands" are 6-8 different few-microsecond
    # kernels per step, so the comparison is per launch, not per class
    dom = max(kt, key=lambda k: kt[k][0] / max(kt[k][1], 1))
    dom_ms, dom_launches, dom_units = kt[dom]
    peak, peak_src = measured_peaks()
    bytes_per_launch = ALGO_BYTES[dom] * dom_units / max(dom_launches, 1)
    us_per_launch = 1000.0 * dom_ms / max(dom_launches, 1)
    achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9 if us_per_launch > 0 else 0.0
    traffic, traffic_src = ncu_traffic(dom, args.workload)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "us_per_launch": us_per_launch,
                "launches_per_step": dom_launches / max(args.profile_steps, 1),
                "share_of_kernel_time": dom_ms / total_kernel_ms,
                "kernel_time_shares": {k: round(v[0] / total_kernel_ms, 4) for k, v in kt.items()}}
    A.close()

    # ------------------------------------------------------------------ end to end, host buffers
    B = fresh_arena()
    hforce = C.c_void_p()
    hstate = C.c_void_p()
    capi.check(lib.b2g_host_alloc(C.byref(hforce), nb * 16))
    capi.check(lib.b2g_host_alloc(C.byref(hstate), nb * 32))
    force_view = np.ctypeslib.as_array(C.cast(hforce, capi.f32p), shape=(nb, 4))
    state_view = np.ctypeslib.as_array(C.cast(hstate, capi.f32p), shape=(nb, 8))
    force_view[:] = 0.0

    def e2e_step():
        B.upload_forces(hforce, 0, nb)                                   # H2D from pinned memory
        if slab_mode:
            B.step(P, None)
            after_step(B)
            capi.check(lib.b2g_download_body_state_async(B.h, 0, nb, hstate))  # D2H into pinned memory
            B.synchronize()
        else:
            # step + D2H of the result into pinned memory; returns when both are complete
            capi.check(lib.b2g_step_download(B.h, C.byref(P), None, 0, nb, hstate))

    for _ in range(args.warmup):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    _, e2e_ms_max, e2e_value = aggregate(nb * args.steps, e2e_ms, device=f"cuda:{local_rank}")
    assert np.isfinite(state_view).all()
    B.close()
    lib.b2g_host_free(hforce)
    lib.b2g_host_free(hstate)

    # ------------------------------------------------------------------ CPU baseline (reference)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.bindings import RefScene
            r = RefScene(scene_name, size, seed)
            r.step(PRESTEP.get(args.workload, 0) + args.warmup)  # same spawn phase as the replicated GPU state
            n = min(args.cpu_sample_steps, args.steps)
            ms = r.time_steps(n)
            # one world on one core: for the batched workloads that is ONE of the arena's worlds, so the
            # unit count is that world's bodies, not the arena's
            cpu_bodies = r.body_count
            cpu = {"value": cpu_bodies * n / (ms / 1000.0), "unit": "body-steps/s", "cores": 1, "kind": "reference",
                   "ms_per_step": ms / n,
                   "sample": f"steps {args.warmup}..{args.warmup + n} of "
                             f"{'one world (' + str(cpu_bodies) + ' bodies) of ' if args.workload in WORLDS_PER_GPU else ''}"
                             f"{args.workload} on 1 host core "
                             "(the reference is single-threaded), oracle/_ref/libb2ref.so compiled from "
                             "/root/reference with -O3 -DNDEBUG"}
        except Exception as exc:  # the oracle is optional for the product, mandatory for the number
            cpu = {"value": None, "unit": "body-steps/s", "cores": 0, "kind": "reference",
                   "sample": f"unavailable: {exc}"}

    if rank == 0:
        line = {
            "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms_max / args.steps,
            "ms_per_step_p50": float(np.percentile(per_step_ms, 50)), "ms_per_step_p99": float(np.percentile(per_step_ms, 99)),
            "higher_is_better": True, "scaling": "strong" if args.workload in SLAB else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "bodies_per_world": nb_world,
                       "worlds": 1 if args.workload in SLAB else per_gpu * world,
                       "halo_bytes_per_step_per_rank": halo_bytes,
                       "worlds_per_gpu": per_gpu, "contacts_mean": float(np.mean(contacts_seen)), "constraints_mean": float(np.mean(constraints_seen)),
                       "colours_max": int(max(colours_seen)), "velocity_iterations": 8, "position_iterations": 3,
                       "dt": 1.0 / 60.0, "sleeping": True, "continuous": False, "solver": "graph-coloured",
                       "l2": "flushed between timed steps (256 MiB memset, outside the event bracket)",
                       "timed_window": f"steps {args.warmup}..{args.warmup + args.steps}"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "body-steps/s", "ms_per_step": e2e_ms_max / args.steps,
                    "h2d_bytes_per_step": nb * 16, "d2h_bytes_per_step": nb * 32},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
