"""Spatial slab decomposition of ONE large world across GPUs (BASELINE config 5).

The world is cut into vertical slabs along x with equal body counts.  Every rank simulates, in
its own arena, the bodies it OWNS plus a GHOST layer: the neighbours' bodies within `halo`
metres of the shared cut plane (and all static bodies).  Boundary contacts are therefore solved
redundantly on both sides; after every step each rank overwrites its ghosts with the owner's
state ("the owner's result wins"), which is the once-per-step halo exchange:

    per neighbour: one packed message of [pos | vel | xf] (48 B) + flags (4 B) per boundary body

sent with NCCL point-to-point (`torch.distributed.batch_isend_irecv`) straight from views of the
arena's device arrays (`b2g_device_views`), or — for the single-process emulation used by the
1-GPU test — copied tensor to tensor.  Ownership and ghost membership are fixed between two
REBALANCES: every K steps (caller's choice) each rank publishes the state of the bodies it owns,
every rank rebuilds the same global picture, the cut planes are moved to equal body counts again and
each rank re-creates its arena from the bodies it now owns plus their ghosts (`rebalance_in_process`,
`rebalance_distributed`).  That is the ownership migration of SURVEY §8e done wholesale instead of
body by body: valid while no body drifts more than `halo` across a cut within K steps (at most 2 m per
step, b2_maxTranslation).  Contacts travel too: every contact is published once (by the owner of its
lowest-numbered movable body) with its manifold and accumulated impulses, and a rebuilt arena starts
from the published contacts whose fixtures it holds, so warm starting survives a rebalance; only the
bodies' sleep timers and the persistent solver colours start over.

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
                              quads=np.concatenate(quads) if quads else np.zeros((1, 4), np.float32),
                              gid=fsel.astype(np.int64)))


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
class SlabRank:
    """one rank's arena + its halo lists on the device (b2g_halo_*: pack / unpack are kernels of the C-ABI
    library on the arena's stream; the transport is libb2cuda_dist.so's NCCL, or — single-process emulation
    — the neighbour arena's send buffer read in place)"""

    def __init__(self, glob, slab, device=0, max_contacts=None):
        from .arena import arena_from_scene
        self.slab = slab
        local = local_scene(glob, slab)
        self.fix_gid = local["fixtures"]["gid"]          # local fixture -> global fixture
        self.fix_body_gid = glob["fixtures"]["body"]     # global fixture -> global body
        self.num_global_fixtures = len(glob["fixtures"]["body"])
        self.body_type = glob["bodies"][:, 11].astype(np.int32)
        self.arena = arena_from_scene(_SceneView(local), max_contacts=max_contacts, device=device)
        self.arena.find_new_contacts()
        lib = self.arena.lib
        for nb in (slab.rank - 1, slab.rank + 1):
            send = np.ascontiguousarray(slab.send_local.get(nb, np.zeros(0, np.int64)), np.int32)
            recv = np.ascontiguousarray(slab.recv_local.get(nb, np.zeros(0, np.int64)), np.int32)
            capi.check(lib.b2g_halo_set_lists(self.arena.h, self.slot_of(nb), len(send), capi.ip(send), len(recv),
                                              capi.ip(recv)), "b2g_halo_set_lists")

    def slot_of(self, nb):
        """C-ABI slot of neighbour rank nb: 0 = lower-x neighbour, 1 = upper-x neighbour"""
        return 0 if nb < self.slab.rank else 1

    def pack(self, nb):
        """enqueues the gather of the state `nb` holds as ghosts; returns (device pointer, bytes)"""
        import ctypes as C
        ptr, nbytes = C.c_void_p(), C.c_int64()
        capi.check(self.arena.lib.b2g_halo_pack(self.arena.h, self.slot_of(nb), C.byref(ptr), C.byref(nbytes)), "b2g_halo_pack")
        return ptr, nbytes.value

    def unpack(self, nb, message=None):
        """enqueues the scatter of a message from `nb` (None: the arena's receive buffer) into the ghosts"""
        capi.check(self.arena.lib.b2g_halo_unpack(self.arena.h, self.slot_of(nb), message), "b2g_halo_unpack")

    def halo_bytes(self):
        return sum(len(ix) * 64 for ix in self.slab.send_local.values())

    def owned_state(self):
        bd = self.arena.download_bodies(what=("pos", "vel", "flags"))
        sl = self.slab.owned_local
        return bd["pos"][sl], bd["vel"][sl], bd["flags"][sl]

    def owned_record(self):
        """[n, 13] float64 rows (global id, pos c.x c.y a, vel v.x v.y w, xf p.x p.y s c, awake, 0) of the
        bodies this rank owns: what a rebalance publishes"""
        bd = self.arena.download_bodies(what=("pos", "vel", "xf", "flags"))
        sl = self.slab.owned_local
        gid = self.slab.global_ids[sl].astype(np.float64)
        awake = ((bd["flags"][sl] & capi.BODY_AWAKE) != 0).astype(np.float64)
        return np.concatenate([gid[:, None], bd["pos"][sl][:, 0:3], bd["vel"][sl][:, 0:3], bd["xf"][sl][:, 0:4],
                               awake[:, None], np.zeros((len(sl), 1))], axis=1)

    def contact_record(self):
        """contacts this rank answers for in a rebalance, by GLOBAL fixture ids, with manifolds and warm-start
        impulses.  A boundary contact lives on both sides of a cut: it is published by the rank that owns the
        contact's non-static body with the smallest global id."""
        c = self.arena.download_contacts()
        ga, gb = self.fix_gid[c["fix_a"]], self.fix_gid[c["fix_b"]]
        owned = np.zeros(len(self.body_type), bool)
        owned[self.slab.global_ids[self.slab.owned_local]] = True
        keep = publishes_contact(ga, gb, self.fix_body_gid, self.body_type, owned)
        return dict(fix_a=ga[keep], fix_b=gb[keep], flags=c["flags"][keep], manifold=c["manifold"][keep],
                    material=c["material"][keep])

    def seed_contacts(self, records, inv_dt0):
        """start this (freshly built) arena from published contacts: those whose two fixtures it holds"""
        inv = np.full(self.num_global_fixtures, -1, np.int64)
        inv[self.fix_gid] = np.arange(len(self.fix_gid))
        fa, fb, fl, mf, mt = [], [], [], [], []
        for r in records:
            if len(r["fix_a"]) == 0:
                continue
            la, lb = inv[r["fix_a"]], inv[r["fix_b"]]
            sel = (la >= 0) & (lb >= 0)
            fa.append(la[sel]); fb.append(lb[sel]); fl.append(r["flags"][sel])
            mf.append(r["manifold"][sel]); mt.append(r["material"][sel])
        if not fa:
            return 0
        fa, fb = np.concatenate(fa).astype(np.int32), np.concatenate(fb).astype(np.int32)
        fl = np.concatenate(fl).astype(np.uint32) & np.uint32(capi.CONTACT_TOUCHING | capi.CONTACT_ENABLED)
        self.arena.upload_contacts(fa, fb, fl, np.concatenate(mf), np.concatenate(mt))
        self.arena.set_inv_dt0(inv_dt0)
        return len(fa)

    def close(self):
        self.arena.close()


class DistTransport:
    """libb2cuda_dist.so: an NCCL communicator of its own (the unique id travels over torch.distributed,
    whatever its backend) for the per-step halo exchange; everything it does is enqueued on the arena's stream"""

    def __init__(self, rank, nranks, device):
        import ctypes as C
        import torch
        import torch.distributed as dist
        self.lib = capi.load_dist()
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            capi.check(self.lib.b2g_dist_unique_id(ident), "b2g_dist_unique_id")
        t = torch.tensor(list(bytes(ident)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.to(torch.device("cuda", device))
        dist.broadcast(t, 0)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        self.h = C.c_void_p()
        capi.check(self.lib.b2g_dist_init(ident, rank, nranks, device, C.byref(self.h)), "b2g_dist_init")

    def exchange(self, sr):
        capi.check(self.lib.b2g_dist_exchange(self.h, sr.arena.h), "b2g_dist_exchange")

    def close(self):
        if self.h:
            self.lib.b2g_dist_destroy(self.h)
            self.h = None


def exchange_distributed(sr, transport):
    """once-per-step halo exchange: pack, ncclSend / ncclRecv per neighbour, unpack — all on the arena's stream"""
    transport.exchange(sr)


def exchange_in_process(ranks):
    """single-process emulation (all slabs on one GPU): every slab packs, then every slab scatters its
    neighbours' send buffers, read in place (the arenas' streams are ordered by a synchronise in between)"""
    msgs = {(sr.slab.rank, nb): sr.pack(nb) for sr in ranks for nb in sr.slab.neighbours}
    for sr in ranks:
        sr.arena.synchronize()
    for sr in ranks:
        for nb in sr.slab.neighbours:
            ptr, _ = msgs[(nb, sr.slab.rank)]
            sr.unpack(nb, ptr)
    for sr in ranks:
        sr.arena.synchronize()


# ------------------------------------------------------------------------------- rebalance
def publishes_contact(fix_a, fix_b, fix_body, body_type, owned):
    """which contacts (pairs of GLOBAL fixture ids) a rank publishes in a rebalance: those whose non-static
    body with the smallest global id it owns.  Every contact has at least one movable body and every
    movable body exactly one owner, so over all ranks each contact is published exactly once."""
    ba, bb = fix_body[fix_a].astype(np.int64), fix_body[fix_b].astype(np.int64)
    big = np.int64(len(body_type))
    lead = np.minimum(np.where(body_type[ba] != capi.STATIC, ba, big), np.where(body_type[bb] != capi.STATIC, bb, big))
    return np.concatenate([owned, [False]])[lead]


def merge_records(glob, records):
    """writes published owner records (SlabRank.owned_record, any order, any number of ranks) into the
    global scene arrays: rows of scene.bodies() = xf(4), c(2), a, v(2), w, awake, type"""
    b = glob["bodies"]
    for rec in records:
        if len(rec) == 0:
            continue
        g = rec[:, 0].astype(np.int64)
        b[g, 4:7] = rec[:, 1:4].astype(np.float32)
        b[g, 7:10] = rec[:, 4:7].astype(np.float32)
        b[g, 0:4] = rec[:, 7:11].astype(np.float32)
        b[g, 10] = rec[:, 11].astype(np.float32)
    return glob


def rebalance_in_process(glob, ranks, halo=3.0, device=0, max_contacts=None, inv_dt0=60.0):
    """single-process emulation: all slabs publish, the world is cut again, every slab is rebuilt.
    Returns (ranks, owner, cuts, migrated) with migrated = bodies that changed owner."""
    nranks = len(ranks)
    old_owner = np.full(len(glob["bodies"]), -1, np.int32)
    for sr in ranks:
        old_owner[sr.slab.global_ids[sr.slab.owned_local]] = sr.slab.rank
    merge_records(glob, [sr.owned_record() for sr in ranks])
    contacts = [sr.contact_record() for sr in ranks]
    for sr in ranks:
        sr.close()
    slabs, owner, cuts = make_slabs(glob, nranks, halo)
    new = [SlabRank(glob, s, device=device, max_contacts=max_contacts) for s in slabs]
    for sr in new:
        sr.seed_contacts(contacts, inv_dt0)
    return new, owner, cuts, int(np.sum(owner != old_owner))


def gather_records(record):
    """all ranks' owner records on every rank (torch.distributed, any backend)"""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, record)
    return out


def rebalance_distributed(glob, sr, halo=3.0, device=0, max_contacts=None, inv_dt0=60.0):
    """one process per GPU: all-gather of the owner records (the only collective of a rebalance; every
    K steps, a few MB), then every rank cuts the same global picture and rebuilds its own slab"""
    import torch.distributed as dist
    nranks = dist.get_world_size()
    records = gather_records(sr.owned_record())
    contacts = gather_records(sr.contact_record())
    old_mine = set(sr.slab.global_ids[sr.slab.owned_local].tolist())
    merge_records(glob, records)
    rank = sr.slab.rank
    sr.close()
    slabs, owner, cuts = make_slabs(glob, nranks, halo)
    new = SlabRank(glob, slabs[rank], device=device, max_contacts=max_contacts)
    new.seed_contacts(contacts, inv_dt0)
    arrived = int(np.sum([g not in old_mine for g in new.slab.global_ids[new.slab.owned_local].tolist()]))
    return new, owner, cuts, arrived


def scene_arrays(scene):
    return dict(bodies=scene.bodies(), params=scene.body_params(), fixtures=scene.fixtures())


def make_slabs(glob, nranks, halo=3.0):
    b = glob["bodies"]
    btype = b[:, 11].astype(np.int32)
    owner, cuts = partition_by_x(b[:, 4].astype(np.float64), btype != capi.STATIC, nranks)
    return [LocalSlab(r, nranks, b[:, 4].astype(np.float64), btype, owner, cuts, halo) for r in range(nranks)], owner, cuts
