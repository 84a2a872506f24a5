"""Spatial slab decomposition of ONE large world across GPUs (BASELINE config 5).

The world is cut into vertical slabs along x with equal body counts.  Every rank simulates, in
its own arena, the bodies it OWNS plus a GHOST layer: the neighbours' bodies within `halo`
metres of the shared cut plane (and all static bodies).  Boundary contacts are therefore solved
redundantly on both sides; after every step each rank overwrites its ghosts with the owner's
state ("the owner's result wins"), which is the once-per-step halo exchange:

    per neighbour: one packed message of [pos | vel | xf] (48 B) + flags (4 B) per boundary body

sent with NCCL point-to-point (`torch.distributed.batch_isend_irecv`) straight from views of the
arena's device arrays (`b2g_device_views`), or — for the single-process emulation used by the
1-GPU test — copied tensor to tensor.  Membership of the ghost layer is fixed at set-up from the
initial positions (round-1 limitation: valid while bodies drift less than `halo` across a cut, as
in a pile settling under gravity); there is no migration yet.

This is an approximation, validated by tolerance against the single-arena result (pile height,
deepest penetration, no lost bodies) — never bit-parity (SURVEY §7 "Hard parts").

Pure-numpy pieces (partition, local scene assembly, exchange lists) have no CUDA dependency and
are covered by CPU tests, including a world_size-2 gloo exchange of the packed messages.
"""
import numpy as np

from . import capi


# ------------------------------------------------------------------------------- partition
def partition_by_x(x, movable, nranks):
    """owner[i] in [0, nranks) for movable bodies (equal counts along x), -1 for static ones;
    cuts[k] = plane between slab k and k+1"""
    owner = np.full(len(x), -1, np.int32)
    idx = np.nonzero(movable)[0]
    order = idx[np.argsort(x[idx], kind="stable")]
    bounds = [int(round(k * len(order) / nranks)) for k in range(nranks + 1)]
    cuts = []
    for r in range(nranks):
        owner[order[bounds[r]:bounds[r + 1]]] = r
        if r + 1 < nranks:
            lo = x[order[bounds[r + 1] - 1]]
            hi = x[order[bounds[r + 1]]]
            cuts.append(0.5 * (float(lo) + float(hi)))
    return owner, np.array(cuts, np.float64)


class LocalSlab:
    """index maps of one rank: which global bodies it holds and what it exchanges with whom"""

    def __init__(self, rank, nranks, x, body_type, owner, cuts, halo):
        self.rank, self.nranks = rank, nranks
        static = np.nonzero(body_type == capi.STATIC)[0]
        owned = np.nonzero(owner == rank)[0]
        ghosts = {}
        sends = {}
        for nb in (rank - 1, rank + 1):
            if nb < 0 or nb >= nranks:
                continue
            cut = cuts[min(rank, nb)]
            near_theirs = np.nonzero((owner == nb) & (np.abs(x - cut) <= halo))[0]
            near_mine = np.nonzero((owner == rank) & (np.abs(x - cut) <= halo))[0]
            ghosts[nb] = near_theirs          # ascending global id: both sides derive the same order
            sends[nb] = near_mine
        self.global_ids = np.concatenate([static, owned] + [ghosts[k] for k in sorted(ghosts)]).astype(np.int64)
        self.local_of = {int(g): i for i, g in enumerate(self.global_ids)}
        self.num_static, self.num_owned = len(static), len(owned)
        self.owned_local = np.arange(len(static), len(static) + len(owned))
        self.send_local = {nb: np.array([self.local_of[int(g)] for g in sends[nb]], np.int64) for nb in sends}
        self.recv_local = {nb: np.array([self.local_of[int(g)] for g in ghosts[nb]], np.int64) for nb in ghosts}
        self.neighbours = sorted(ghosts)


def local_scene(glob, slab):
    """cuts the global scene arrays (GpuScene/RefScene .bodies(), .body_params(), .fixtures()) down
    to one rank's bodies; body and shape indices are remapped"""
    b, p, fx = glob["bodies"], glob["params"], glob["fixtures"]
    gids = slab.global_ids
    keep_body = np.zeros(len(b), bool)
    keep_body[gids] = True
    fsel = np.nonzero(keep_body[fx["body"]])[0]
    remap = np.full(len(b), -1, np.int64)
    remap[gids] = np.arange(len(gids))
    # shape pool: gather each kept fixture's record (vectorised: a 1 M-body world has 1 M records)
    off = fx["shape_off"][fsel].astype(np.int64)
    ftype = fx["type"][fsel]
    sizes = np.where(ftype == 0, 1, np.where(ftype == 1, 3, 1 + fx["quads"][off, 3].astype(np.int64)))
    new_off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if len(sizes) else np.zeros(0, np.int64)
    total = int(sizes.sum())
    src = np.repeat(off - new_off, sizes) + np.arange(total)
    quads = [fx["quads"][src]] if total else []
    offs = new_off
    return dict(bodies=b[gids], params=p[gids],
                fixtures=dict(body=remap[fx["body"][fsel]].astype(np.int32), type=fx["type"][fsel],
                              shape_off=np.array(offs, np.int32), filter=fx["filter"][fsel],
                              material=fx["material"][fsel], sensor=fx["sensor"][fsel],
                              quads=np.concatenate(quads) if quads else np.zeros((1, 4), np.float32)))


class _SceneView:
    """duck-types the scene interface arena_from_scene() reads"""

    def __init__(self, d):
        self.d = d

    def bodies(self):
        return self.d["bodies"]

    def body_params(self):
        return self.d["params"]

    def fixtures(self):
        return self.d["fixtures"]

    def joints(self):
        return dict(bodies=np.zeros((0, 2), np.int32), anchors=np.zeros((0, 4), np.float32),
                    params=np.zeros((0, 12), np.float32))


# --------------------------------------------------------------------------------- device side
class _CudaView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


class SlabRank:
    """one rank's arena + halo pack / unpack on the device (torch views of the arena arrays)"""

    def __init__(self, glob, slab, device=0, max_contacts=None):
        import torch
        from .arena import arena_from_scene
        self.slab = slab
        self.torch = torch
        self.arena = arena_from_scene(_SceneView(local_scene(glob, slab)), max_contacts=max_contacts, device=device)
        self.arena.find_new_contacts()
        v = self.arena.device_views()
        dev = torch.device("cuda", device)
        cap = v.capacity
        self.t_pos = torch.as_tensor(_CudaView(v.pos, (cap, 4), "<f4"), device=dev)
        self.t_vel = torch.as_tensor(_CudaView(v.vel, (cap, 4), "<f4"), device=dev)
        self.t_xf = torch.as_tensor(_CudaView(v.xf, (cap, 4), "<f4"), device=dev)
        self.t_flags = torch.as_tensor(_CudaView(v.flags, (cap,), "<i4"), device=dev)
        self.send_idx = {nb: torch.as_tensor(ix, device=dev) for nb, ix in slab.send_local.items()}
        self.recv_idx = {nb: torch.as_tensor(ix, device=dev) for nb, ix in slab.recv_local.items()}
        self.recv_buf = {nb: torch.empty((len(ix), 13), dtype=torch.float32, device=dev)
                         for nb, ix in slab.recv_local.items()}

    def pack(self, nb):
        """[n, 13] float32: pos(4) vel(4) xf(4) flags-as-float-bits(1) of the bodies `nb` holds as ghosts"""
        ix = self.send_idx[nb]
        t = self.torch
        return t.cat([self.t_pos[ix], self.t_vel[ix], self.t_xf[ix],
                      self.t_flags[ix].view(t.float32).unsqueeze(1)], dim=1).contiguous()

    def unpack(self, nb, msg):
        ix = self.recv_idx[nb]
        self.t_pos[ix] = msg[:, 0:4]
        self.t_vel[ix] = msg[:, 4:8]
        self.t_xf[ix] = msg[:, 8:12]
        self.t_flags[ix] = msg[:, 12].contiguous().view(self.torch.int32)

    def halo_bytes(self):
        return sum(len(ix) * 52 for ix in self.slab.send_local.values())

    def owned_state(self):
        bd = self.arena.download_bodies(what=("pos", "vel", "flags"))
        sl = self.slab.owned_local
        return bd["pos"][sl], bd["vel"][sl], bd["flags"][sl]


def exchange_distributed(sr):
    """once-per-step halo exchange over NCCL point-to-point (one message per neighbour)"""
    import torch.distributed as dist
    ops, outs = [], {}
    for nb in sr.slab.neighbours:
        outs[nb] = sr.pack(nb)
        ops.append(dist.P2POp(dist.isend, outs[nb], nb))
        ops.append(dist.P2POp(dist.irecv, sr.recv_buf[nb], nb))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for nb in sr.slab.neighbours:
        sr.unpack(nb, sr.recv_buf[nb])


def exchange_in_process(ranks):
    """single-process emulation (all slabs on one GPU): the same pack / unpack, tensor to tensor"""
    msgs = {(sr.slab.rank, nb): sr.pack(nb) for sr in ranks for nb in sr.slab.neighbours}
    for sr in ranks:
        for nb in sr.slab.neighbours:
            sr.unpack(nb, msgs[(nb, sr.slab.rank)])


def scene_arrays(scene):
    return dict(bodies=scene.bodies(), params=scene.body_params(), fixtures=scene.fixtures())


def make_slabs(glob, nranks, halo=3.0):
    b = glob["bodies"]
    btype = b[:, 11].astype(np.int32)
    owner, cuts = partition_by_x(b[:, 4].astype(np.float64), btype != capi.STATIC, nranks)
    return [LocalSlab(r, nranks, b[:, 4].astype(np.float64), btype, owner, cuts, halo) for r in range(nranks)], owner, cuts
