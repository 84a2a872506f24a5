#!/usr/bin/env python
"""bench.py — ms per b2World::Step and body-steps/s (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload NAME]

A "step" is one b2World::Step of the workload.  Defaults (`--workload auto`):
  N = 1 : mixed_100k — BASELINE.json configs[2], the 100k-body scene the north_star's >= 50x target is
          quoted on (configs[1] many_pyramids, configs[3..4] are `--workload` choices).
  N > 1 : tumbler_worlds — configs[3]: batched independent tumbler worlds, 1024 per GPU, sharded over
          ranks with NO data-path collective (weak scaling); every world is a different world (16 spawn
          variants x a per-world velocity perturbation), not a clone.

Every workload is PRE-ROLLED outside the timed region to its all-in-contact window, in both arms
(`config.preroll_steps`): the timed window is steps preroll+W .. preroll+W+K of the scene whatever W and K
are, so a short driver run times the same physics as a long one.

  value  : device-resident throughput — bodies x K / sum of per-step CUDA-event times on the
           arena's stream; L2 is flushed (256 MiB memset) between timed steps.
  e2e    : the same metric through the C-ABI with HOST buffers on the same step window: every step
           uploads the force accumulators from pinned host memory (b2g_upload_forces), steps and reads
           every body's transform + velocity back to pinned host memory (b2g_step_download), all inside
           the timed region, one step at a time (no pipelining across steps).
  roofline     : the dominant KERNEL (largest CUDA-event time per launch; the library times every launch
                 by kernel class) measured on the SAME step window of an identical, deterministic arena;
                 algorithmic bytes from SURVEY.md §8(d) (table in DESIGN.md) / measured launch time,
                 against MEASURED_PEAKS.json's HBM copy bandwidth; `traffic` = DRAM bytes per launch of
                 that kernel from the newest committed ncu --set full summary of this workload.
  cpu_baseline : the reference's own CPU Step (oracle/_ref, compiled from /root/reference) on a
                 bounded sample of the same window, rank 0 / N=1 only.
`--impl reference` times that CPU implementation alone (one world per thread).
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
    "query": 2300.0,
}
# kernels whose launch covers several of §8(d)'s passes: bytes per constraint follow the iteration counts
SOLVE_PARTS = {
    # oversize islands: constraint preparation is a separate launch (k_prepare, class "prepare")
    "big_solve": lambda vi, pi: 180.0 + vi * 196.0 + pi * 136.0 + 32.0,
    # tiled oversize islands: everything of b2Island::Solve in one launch
    "big_tiles": lambda vi, pi: 292.0 + 180.0 + vi * 196.0 + pi * 136.0 + 32.0,
}
VEL_ITERS, POS_ITERS = 8, 3

# name: scene, size, seed, preroll (steps of the scene before the timed window, outside the timed region),
#       per_gpu (independent worlds per GPU), variants (different worlds per arena), host_prestep (steps made
#       through the drop-in API before the state is copied: the tumbler's spawn phase)
WORKLOADS = {
    "mixed_100k": dict(scene="mixed", size=100000, seed=12345, preroll=300,
                       desc="config 3: 100k circles + convex polygons settling into a container, sleeping on"),
    "many_pyramids": dict(scene="many_pyramids", size=100, seed=0, preroll=60,
                          desc="config 2: 100 independent 20-row pyramids (21001 bodies) in one world"),
    "pyramid": dict(scene="pyramid", size=20, seed=0, preroll=60, desc="config 1: testbed pyramid, 20 rows (211 bodies)"),
    "mixed_10k": dict(scene="mixed", size=10000, seed=12345, preroll=150,
                      desc="10k circles + convex polygons settling into a container"),
    # batched independent worlds in ONE arena per GPU (world id in the broadphase key), sharded
    # round-robin over ranks
    "pyramid_worlds": dict(scene="pyramid", size=20, seed=0, preroll=60, per_gpu=512,
                           desc="512 independent 20-row pyramid worlds per GPU (108k bodies), one arena"),
    # BASELINE config 4: motor-driven tumbler (benchmarks.h b3) with 500 boxes per world.  The spawn phase
    # (one box per step) runs through the drop-in API for each of the 16 spawn variants; the arena then
    # holds worlds_per_gpu worlds (variant w mod 16 + a per-world velocity perturbation) and is pre-rolled
    "tumbler_worlds": dict(scene="tumbler", size=500, seed=0, preroll=120, per_gpu=1024, variants=16, host_prestep=520,
                           desc="config 4: 1024 independent 500-box tumbler worlds per GPU (8192 on 8 GPUs; "
                                "514k bodies per arena), every world different"),
    # BASELINE config 5: ONE world cut into x-slabs, one per GPU, ghost layer refreshed by a
    # per-step NCCL halo exchange (strong scaling: the world is fixed, the slab shrinks with N)
    "mixed_slab_400k": dict(scene="mixed", size=400000, seed=12345, preroll=300, slab=True,
                            desc="one 400k-body world, x-slab per GPU, per-step NCCL halo exchange"),
    "mixed_slab_1m": dict(scene="mixed", size=1000000, seed=12345, preroll=300, slab=True,
                          desc="config 5: one 1M-body world, x-slab per GPU, per-step NCCL halo exchange"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--worlds-per-gpu", type=int, default=0,
                    help="batched-world workloads: independent worlds per GPU (config 4 = 8192 worlds / 8 GPUs = 1024)")
    ap.add_argument("--preroll", type=int, default=-1, help="override the workload's pre-roll (steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end leg")
    ap.add_argument("--no-roofline", action="store_true", help="profiling runs: skip the per-kernel pass")
    ap.add_argument("--cpu-sample-steps", type=int, default=40)
    ap.add_argument("--profiler-range", action="store_true",
                    help="profiling runs: cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    return ap.parse_args()


def pick_workload(args, world):
    name = args.workload
    if name == "auto":
        name = "mixed_100k" if max(world, args.gpus) <= 1 else "tumbler_worlds"
    W = dict(WORKLOADS[name])
    W["name"] = name
    W.setdefault("per_gpu", 1)
    W.setdefault("variants", 1)
    W.setdefault("host_prestep", 0)
    W.setdefault("slab", False)
    if args.worlds_per_gpu > 0 and WORKLOADS[name].get("per_gpu"):
        W["per_gpu"] = args.worlds_per_gpu
    if args.preroll >= 0:
        W["preroll"] = args.preroll
    W["batched"] = "per_gpu" in WORKLOADS[name]
    return W


def variant_seed(W, v):
    """scene seed of spawn variant v (variants = 1: the workload's own seed)"""
    return W["seed"] if W["variants"] <= 1 else v + 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clocks and throttle reasons with nvidia-smi while the timed region runs (rank 0 only)"""

    def __init__(self, device, enabled=True):
        super().__init__(daemon=True)
        self.device = device
        self.enabled = enabled
        self.stop_flag = threading.Event()
        self.samples = []
        self.reasons = set()
        self.sm_max = None

    def sample(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
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

    def run(self):
        while self.enabled and not self.stop_flag.is_set():
            self.sample()
            self.stop_flag.wait(0.1)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "window": "pre-roll + warm-up + timed steps (continuous load)"}


def window_text(W, warmup, steps):
    a = W["host_prestep"] + W["preroll"] + warmup
    return (f"steps {a}..{a + steps} of the scene ({W['host_prestep'] + W['preroll']} pre-rolled outside the timed "
            f"region + {warmup} warm-up)")


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU b2World::Step, one world per host thread."""
    if rank != 0:
        return
    from oracle.bindings import RefScene
    W = pick_workload(args, world)
    # one world per GPU of our arm for the single-world workloads; for the batched workloads a
    # bounded sample of one world per host thread (the per-world cost is the same)
    nworlds = max(1, args.gpus)
    ncpu = os.cpu_count() or 1
    if W["batched"]:
        nworlds = min(ncpu, W["per_gpu"] * max(1, args.gpus))
    nworlds = min(nworlds, max(1, ncpu))
    worlds = [RefScene(W["scene"], W["size"], variant_seed(W, i % W["variants"])) for i in range(nworlds)]
    nb = [0] * nworlds
    times = [0.0] * nworlds

    def work(i):
        # pre-roll to the same window as our arm, outside the timed region
        worlds[i].step(W["host_prestep"] + W["preroll"] + args.warmup)
        nb[i] = worlds[i].body_count
        times[i] = worlds[i].time_steps(args.steps)

    th = [threading.Thread(target=work, args=(i,)) for i in range(nworlds)]
    [t.start() for t in th]
    [t.join() for t in th]
    ms = max(times)
    value = sum(nb) * args.steps / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if W["slab"] else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": W["name"], "description": W["desc"], "bodies_per_world": nb[0], "worlds": nworlds,
                   "velocity_iterations": VEL_ITERS, "position_iterations": POS_ITERS, "dt": 1.0 / 60.0, "sleeping": True,
                   "continuous": False, "preroll_steps": W["host_prestep"] + W["preroll"],
                   "timed_window": window_text(W, args.warmup, args.steps),
                   "contacts_end": int(worlds[0].contact_count)},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": nworlds, "kind": "reference",
                         "sample": f"{window_text(W, args.warmup, args.steps)} of {W['name']}, "
                                   f"{nworlds} world(s), one thread each (the reference is single-threaded)"},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


KERNEL_OF_CLASS = {"fused_solve": "k_solve_bins_fused", "bp_traverse": "k_bp_traverse", "narrowphase": "k_narrowphase",
                   "big_solve": "k_big_solve", "big_tiles": "k_big_tiles"}


def ncu_traffic(cls, workload):
    """DRAM bytes per launch of the kernel, from the newest committed ncu capture of THIS workload
    (profiles/*_dram_traffic_<workload>.csv: kernel, us, GB/s, MB per launch — dram__bytes.sum.per_second x
    gpu__time_duration of a launch in the timed window); None when there is none."""
    import glob
    if cls not in KERNEL_OF_CLASS:
        return None, None
    for f in reversed(sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_dram_traffic_{workload}.csv")))):
        try:
            for line in open(f):
                cells = line.strip().split(",")
                if len(cells) == 4 and cells[0].split("<")[0] == KERNEL_OF_CLASS[cls]:
                    return float(cells[3]) * 1e6, os.path.basename(f)
        except (OSError, ValueError):
            continue
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

    W = pick_workload(args, world)
    lib = capi.load_cuda()
    # host-side scene construction through the drop-in C++ API (the tumbler's spawn phase is stepped there)
    scenes = []
    for v in range(W["variants"]):
        s = GpuScene(W["scene"], W["size"], variant_seed(W, v), device=local_rank)
        if W["host_prestep"]:
            s.step(W["host_prestep"])
        scenes.append(s)
    nb_world = scenes[0].body_count

    from box2d_optimized_b200.sharding import aggregate, shard_worlds
    per_gpu = W["per_gpu"]
    my_worlds = shard_worlds(per_gpu * world, rank, world)   # independent worlds, no data-path collective
    copies = len(my_worlds)
    nb = nb_world * copies
    cap_contacts = max(4096, 8 * nb)

    slab_mode = W["slab"] and world > 1
    halo_bytes = 0
    if slab_mode:
        from box2d_optimized_b200.slab import DistTransport, SlabRank, make_slabs, scene_arrays
        transport = DistTransport(rank, world, local_rank)   # libb2cuda_dist.so: NCCL on the arena's stream
        glob = scene_arrays(scenes[0])
        slabs, _, _ = make_slabs(glob, world, halo=3.0)
        nb = slabs[rank].num_owned  # bodies this rank advances (ghosts are redundant work)
        copies = 1

    def perturb(A):
        """batched worlds: a per-world velocity perturbation (1 mm/s) keyed by the GLOBAL world id, so no two
        worlds of the job evolve in lock-step (the pre-roll amplifies it: a tumbler is chaotic)"""
        if not W["batched"] or copies <= 1:
            return
        d = A.download_bodies(what=("vel", "flags"))
        vel, fl = d["vel"], d["flags"]
        dyn = ((fl >> capi.BODY_TYPE_SHIFT) & 3) == capi.DYNAMIC
        for k, wid in enumerate(my_worlds):
            rng = np.random.default_rng(1000003 * (wid + 1))
            sl = slice(k * nb_world, (k + 1) * nb_world)
            noise = rng.standard_normal((nb_world, 3)).astype(np.float32) * np.float32(1e-3)
            vel[sl, :3] += noise * dyn[sl, None]
        A.upload_bodies(0, vel=vel)

    P = Arena.params(vel_iters=VEL_ITERS, pos_iters=POS_ITERS)
    st = capi.StepStats()

    def after_step(A):
        # the once-per-step halo exchange of the slab decomposition: pack kernel, ncclSend / ncclRecv per
        # neighbour, unpack kernel, all enqueued on the arena's stream behind the step (no host wait)
        if slab_mode:
            transport.exchange(A._slab)

    def fresh_arena(before_stepping=None):
        """an arena at the start of the timed window: built from the scene(s), perturbed, pre-rolled and warmed
        up — all outside the timed region.  Deterministic: every arena built here reaches the same state."""
        if slab_mode:
            sr = SlabRank(glob, slabs[rank], device=local_rank, max_contacts=max(4096, 8 * len(slabs[rank].global_ids)))
            sr.arena._slab = sr
            A = sr.arena
        else:
            A = arena_from_scene(scenes, max_contacts=cap_contacts, device=local_rank, copies=copies, num_worlds=copies)
            perturb(A)
            A.find_new_contacts()
        if before_stepping is not None:
            before_stepping()
        for _ in range(W["preroll"] + args.warmup):
            A.step(P, None)
            after_step(A)
        A.synchronize()
        return A

    # ------------------------------------------------------------------ device-resident value
    # clocks and throttle reasons are sampled from here on: the pre-roll and warm-up are the same kernels on the same
    # GPU right before the timed steps, and a 20-step timed region alone (tens of milliseconds) is shorter than one
    # nvidia-smi query
    sampler = ClockSampler(local_rank, enabled=(rank == 0))
    A = fresh_arena(before_stepping=sampler.start)
    ext = torch.cuda.ExternalStream(A.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    if slab_mode:
        halo_bytes = A._slab.halo_bytes()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches = 0
    contacts_seen, constraints_seen, colours_seen, awake_seen, overflow_seen, rounds_seen = [], [], [], [], [], []
    if args.profiler_range:
        torch.cuda.profiler.start()
    for k in range(args.steps):
        with torch.cuda.stream(ext):
            flush.zero_()  # evict the step's working set from L2 (outside the timed bracket)
        starts[k].record(ext)
        A.step(P, st)
        after_step(A)
        ends[k].record(ext)
        launches += st.num_launches
        contacts_seen.append(st.num_contacts)
        constraints_seen.append(st.num_constraints)
        colours_seen.append(st.num_colours)
        awake_seen.append(st.num_awake)
        overflow_seen.append(st.num_overflow)
        rounds_seen.append(st.colour_rounds)
    A.synchronize()
    torch.cuda.synchronize()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    if rank == 0 and not sampler.samples:
        sampler.sample()
    clocks = sampler.result()
    per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    elapsed_ms = sum(per_step_ms)
    _, elapsed_ms_max, value = aggregate(nb * args.steps, elapsed_ms, device=f"cuda:{local_rank}")
    A.close()

    # ------------------------------------------------------------------ per-kernel roofline pass
    # An identical arena (same scene, same pre-roll: runs are bit-reproducible) stepped over the SAME window
    # with every launch bracketed by CUDA events on the arena's stream.
    roofline = None
    if not args.no_roofline:
        R = fresh_arena()
        R.set_kernel_timing(True)
        for _ in range(args.steps):
            R.step(P, st)
            after_step(R)
        kt = R.kernel_timing()
        R.set_kernel_timing(False)
        R.close()
        kt = {k: v for k, v in kt.items() if v[1] > 0}
        total_kernel_ms = sum(v[0] for v in kt.values()) or 1.0
        # the dominant KERNEL: classes such as "colour" or "islands" are several few-microsecond
        # kernels per step, so the comparison is per launch, not per class
        dom = max(kt, key=lambda k: kt[k][0] / max(kt[k][1], 1))
        peak, peak_src = measured_peaks()

        def kernel_line(cls):
            ms, n, units = kt[cls]
            per_unit = SOLVE_PARTS[cls](VEL_ITERS, POS_ITERS) if cls in SOLVE_PARTS else ALGO_BYTES[cls]
            bpl = per_unit * units / max(n, 1)
            us = 1000.0 * ms / max(n, 1)
            ach = bpl / (us * 1e-6) / 1e9 if us > 0 else 0.0
            traffic, src = ncu_traffic(cls, W["name"])
            return {"kernel": KERNEL_OF_CLASS.get(cls, cls), "class": cls, "achieved": ach, "frac": ach / peak,
                    "algorithmic_bytes_per_launch": bpl, "bytes_per_unit": per_unit, "units_per_launch": units / max(n, 1),
                    "us_per_launch": us, "launches_per_step": n / max(args.steps, 1),
                    "share_of_kernel_time": ms / total_kernel_ms, "traffic": traffic, "traffic_source": src,
                    "frac_dram": (traffic / (us * 1e-6) / 1e9 / peak) if (traffic and us > 0) else None}

        d = kernel_line(dom)
        roofline = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": d["frac"], "traffic": d["traffic"], "traffic_source": d["traffic_source"],
                    "frac_dram": d["frac_dram"], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                    "bytes_per_unit": d["bytes_per_unit"], "units_per_launch": d["units_per_launch"],
                    "us_per_launch": d["us_per_launch"], "launches_per_step": d["launches_per_step"],
                    "share_of_kernel_time": d["share_of_kernel_time"],
                    "window": "same steps as the timed window, on an identical arena",
                    "kernel_time_shares": {k: round(v[0] / total_kernel_ms, 4) for k, v in kt.items()},
                    "kernel_us_per_step": {k: round(1000.0 * v[0] / max(args.steps, 1), 2) for k, v in kt.items()},
                    "solver_kernels": [kernel_line(c) for c in ("big_tiles", "big_solve", "fused_solve") if c in kt]}

    # ------------------------------------------------------------------ end to end, host buffers
    e2e = None
    if not args.no_e2e:
        B = fresh_arena()
        hforce = C.c_void_p()
        hstate = C.c_void_p()
        nall = B.num_bodies if slab_mode else nb
        capi.check(lib.b2g_host_alloc(C.byref(hforce), nall * 16))
        capi.check(lib.b2g_host_alloc(C.byref(hstate), nall * 32))
        force_view = np.ctypeslib.as_array(C.cast(hforce, capi.f32p), shape=(nall, 4))
        state_view = np.ctypeslib.as_array(C.cast(hstate, capi.f32p), shape=(nall, 8))
        force_view[:] = 0.0

        def e2e_step():
            B.upload_forces(hforce, 0, nall)                                   # H2D from pinned memory
            if slab_mode:
                B.step(P, None)
                after_step(B)
                capi.check(lib.b2g_download_body_state_async(B.h, 0, nall, hstate))  # D2H into pinned memory
                B.synchronize()
            else:
                # step + D2H of the result into pinned memory; returns when both are complete
                capi.check(lib.b2g_step_download(B.h, C.byref(P), None, 0, nall, hstate))

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
        e2e = {"value": e2e_value, "unit": "body-steps/s", "ms_per_step": e2e_ms_max / args.steps,
               "h2d_bytes_per_step": nall * 16, "d2h_bytes_per_step": nall * 32}

    # ------------------------------------------------------------------ CPU baseline (reference)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.bindings import RefScene
            r = RefScene(W["scene"], W["size"], variant_seed(W, 0))
            r.step(W["host_prestep"] + W["preroll"] + args.warmup)  # same window, outside the timed sample
            n = max(1, min(args.cpu_sample_steps, args.steps))
            ms = r.time_steps(n)
            # one world on one core: for the batched workloads that is ONE of the arena's worlds, so the
            # unit count is that world's bodies, not the arena's
            cpu_bodies = r.body_count
            cpu = {"value": cpu_bodies * n / (ms / 1000.0), "unit": "body-steps/s", "cores": 1, "kind": "reference",
                   "ms_per_step": ms / n, "contacts_end": int(r.contact_count),
                   "sample": f"first {n} steps of the timed window ({window_text(W, args.warmup, n)}) of "
                             f"{'one world (' + str(cpu_bodies) + ' bodies) of ' if W['batched'] else ''}"
                             f"{W['name']} on 1 host core "
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
            "higher_is_better": True, "scaling": "strong" if W["slab"] else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": W["name"], "description": W["desc"], "bodies_per_world": nb_world,
                       "worlds": 1 if W["slab"] else per_gpu * world,
                       "halo_bytes_per_step_per_rank": halo_bytes,
                       "worlds_per_gpu": per_gpu, "world_variants": W["variants"],
                       "contacts_mean": float(np.mean(contacts_seen)), "constraints_mean": float(np.mean(constraints_seen)),
                       "awake_bodies_mean": float(np.mean(awake_seen)),
                       "colours_max": int(max(colours_seen)), "serial_bucket_constraints_mean": float(np.mean(overflow_seen)),
                       "colour_rounds_mean": float(np.mean(rounds_seen)), "velocity_iterations": VEL_ITERS, "position_iterations": POS_ITERS,
                       "dt": 1.0 / 60.0, "sleeping": True, "continuous": False, "solver": "graph-coloured",
                       "l2": "flushed between timed steps (256 MiB memset, outside the event bracket)",
                       "preroll_steps": W["host_prestep"] + W["preroll"],
                       "timed_window": window_text(W, args.warmup, args.steps)},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
