"""Array-level pipeline over the C-ABI: events -> store -> X -> W -> top-k, all device resident.

This is the layer ``UserItemInteractions`` / ``SLIM`` sit on, exposed so that callers who already
hold event columns on the device (bench.py, the multi-GPU driver) can run the same kernels
without the host object model.  Sharding helpers for ``torch.distributed`` live here too: X is
replicated; the fit is partitioned by item column -- by default every rank completes only the Gram rows of
its own targets out of peer memory (``fit_owner_rows``), alternatively the whole triangle is exchanged
(``gram_sharded`` + ``fit_sharded``); scoring is partitioned by query user (``recommend_query_sharded``, lists
all-gathered) or by item column (``recommend_sharded``, per-rank top-k lists merged by ``rt_topk_merge``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from . import device as D
from ._lib import FitConfig


@dataclass
class StoreState:
    keys: object      # int64 device (bit pattern of u64 user<<32|item), sorted
    vals: object      # float64 device
    stamps: object    # float64 device
    n_pairs: int
    max_ts: float
    max_user: int
    max_item: int


def empty_store() -> StoreState:
    return StoreState(None, None, None, 0, 0.0, 0, 0)


def fold_events(state: StoreState, du, di, dts, dd, *, upsert: bool, min_value: float, max_value: float,
                decay_rate: Optional[float]) -> StoreState:
    """K1: fold a batch of device-resident events (int32, int32, f64, f64) into the store."""
    t = D.require_cuda()
    n = int(du.numel())
    cap = state.n_pairs + n
    ok = D.empty(max(cap, 1), t.int64); ov = D.empty(max(cap, 1), t.float64); os_ = D.empty(max(cap, 1), t.float64)
    n_out, mts, mu, mi = C.c_int64(0), C.c_double(0), C.c_int32(0), C.c_int32(0)
    rate = decay_rate if decay_rate is not None else float("nan")
    _lib.check(_lib.load().rt_store_fold(D.ptr(du), D.ptr(di), D.ptr(dts), D.ptr(dd), n, int(bool(upsert)),
                                         float(min_value), float(max_value), rate, D.ptr(state.keys), D.ptr(state.vals),
                                         D.ptr(state.stamps), state.n_pairs, float(state.max_ts), int(state.max_user),
                                         int(state.max_item), D.ptr(ok), D.ptr(ov), D.ptr(os_), cap, C.byref(n_out),
                                         C.byref(mts), C.byref(mu), C.byref(mi), D.stream_ptr()), "rt_store_fold")
    p = int(n_out.value)
    return StoreState(ok[:p], ov[:p], os_[:p], p, float(mts.value), int(mu.value), int(mi.value))


def build_matrix(state: StoreState, *, decay_rate: Optional[float], max_ts: Optional[float] = None,
                 n_users: Optional[int] = None, n_items: Optional[int] = None, item_mask=None) -> D.DeviceMatrix:
    """K2: float32 CSR + CSC (+ COO column ids) of the store, decay applied at ``max_ts``."""
    t = D.require_cuda()
    n = state.n_pairs
    n_users = state.max_user + 1 if n_users is None else n_users
    n_items = state.max_item + 1 if n_items is None else n_items
    max_ts = state.max_ts if max_ts is None else max_ts
    rptr = D.empty(n_users + 1, t.int32); ridx = D.empty(max(n, 1), t.int32); rval = D.empty(max(n, 1), t.float32)
    cptr = D.empty(n_items + 1, t.int32); cidx = D.empty(max(n, 1), t.int32); cval = D.empty(max(n, 1), t.float32)
    ccol = D.empty(max(n, 1), t.int32)
    nnz, nonneg = C.c_int64(0), C.c_int(0)
    rate = decay_rate if decay_rate is not None else float("nan")
    _lib.check(_lib.load().rt_store_build(D.ptr(state.keys), D.ptr(state.vals), D.ptr(state.stamps), n, rate, float(max_ts),
                                          n_users, n_items, D.ptr(item_mask), D.ptr(rptr), D.ptr(ridx), D.ptr(rval),
                                          D.ptr(cptr), D.ptr(cidx), D.ptr(cval), D.ptr(ccol), C.byref(nnz),
                                          C.byref(nonneg), D.stream_ptr()), "rt_store_build")
    return D.DeviceMatrix(n_users, n_items, int(nnz.value), rptr, ridx, rval, cptr, cidx, cval, ccol, bool(nonneg.value))


def touched_items(di, dd, n_items: int):
    """Item bookkeeping of a device-resident event batch (rt_events_item_stats32; base.py:85-94 records the item of
    every event for the partial fit, slim.py:31-53): returns ``(mask uint8 [n_items] device, targets int32 device)``
    -- the 0/1 column mask ``build_matrix(item_mask=...)`` takes and the ascending list of touched item ids."""
    t = D.require_cuda()
    n = int(di.numel())
    cnt = D.empty(n_items, t.int32); last = D.empty(n_items, t.int32); seen = D.empty(n_items, t.uint8)
    _lib.check(_lib.load().rt_events_item_stats32(D.ptr(di), D.ptr(dd), n, int(n_items), D.ptr(cnt), D.ptr(last),
                                                  D.ptr(seen), D.stream_ptr()), "rt_events_item_stats32")
    ids = np.flatnonzero(seen.cpu().numpy()).astype(np.int32)   # a few KB; the host keeps this list anyway (slim.py:29)
    return seen, D.to_dev(ids)


# ------------------------------------------------------------------------------------------ sharding
def item_shard(n_items: int, rank: int, world: int, cptr_host: Optional[np.ndarray] = None) -> Tuple[int, int]:
    """Contiguous item range [j0, j1) owned by ``rank``: equal column counts.  One ElasticNet solve costs
    about the same for every target (nn candidates each), and the scoring work of a shard is the number of
    W entries in its columns (<= nn per column), so the count -- not the stored entries of X, which would
    hand the popular head to one rank and nearly every column to the last -- is what balances the ranks.
    ``cptr_host`` is accepted for compatibility and ignored."""
    if world <= 1:
        return 0, n_items
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def item_stride(n_items: int, rank: int, world: int):
    """Interleaved target columns ``rank, rank + world, ...`` (device int32).  The cost of one solve grows
    with the popularity of the target (more live candidates), and real item ids are often correlated with
    popularity or age, so a stride balances the fit better than contiguous ranges.  Used when the scoring
    side does not need contiguous column shards (query-partitioned scoring)."""
    t = D.require_cuda()
    return t.arange(rank, n_items, max(world, 1), dtype=t.int32, device=D.dev())


def query_cuts(rptr, users, world: int) -> list:
    """Cut points (len world + 1) of the query list, balanced by the stored entries of the queried rows
    (the scoring work of a user grows with its row length).  Computed from replicated data, so every rank
    derives the same list."""
    t = D.torch()
    Q = int(users.numel())
    if world <= 1 or Q == 0:
        return [0] + [Q] * max(world, 1)
    ul = users.long()
    work = (rptr[ul + 1] - rptr[ul]).long() + 8  # + fixed per-user cost
    csum = t.cumsum(work, 0)
    frac = t.arange(1, world, dtype=t.int64, device=csum.device)
    want = (csum[-1] * frac) // world
    cuts = [0] + t.searchsorted(csum, want).tolist() + [Q]
    for r in range(1, world + 1):
        cuts[r] = min(max(cuts[r], cuts[r - 1]), Q)
    return cuts


def query_shard(rptr, users, rank: int, world: int) -> Tuple[int, int]:
    cuts = query_cuts(rptr, users, world)
    return cuts[rank], cuts[rank + 1]


def exchange_slabs(M, cuts, group=None):
    """All-gather of unequal row slabs: rank r owns rows [cuts[r], cuts[r+1]) of ``M`` (same cuts on every
    rank); afterwards every rank holds all rows.  One broadcast per non-empty slab, issued asynchronously."""
    import torch.distributed as dist
    works = []
    for r in range(len(cuts) - 1):
        a, b = int(cuts[r]), int(cuts[r + 1])
        if b > a:
            works.append(dist.broadcast(M[a:b], src=dist.get_global_rank(group, r) if group is not None else r,
                                        group=group, async_op=True))
    for w in works:
        w.wait()


class PeerSlabs:
    """One I x I float32 slab buffer per rank, mapped into every rank of the node through CUDA IPC, plus the
    stream-ordered barriers the fused exchange needs.  Buffers are cached per (group, n_items) and reused."""

    _cache: dict = {}

    def __init__(self, n_items: int, rank: int, world: int, group=None):
        import torch.distributed as dist
        t = D.require_cuda()
        lib = _lib.load()
        self.n_items, self.rank, self.world, self.group = n_items, rank, world, group
        self.own, self.ptrs, self.error, self.rows_alloc = 0, [0] * world, None, 0
        self._flag = t.zeros(1, dtype=t.int32, device=D.dev())
        # every rank takes part in both collectives below whatever happens locally, so a failure on one
        # rank (IPC not permitted in this container, out of memory) turns into an agreed fallback, not a hang
        payload = None
        try:
            own = C.c_void_p(0)
            handle = (C.c_uint8 * 64)()
            # room for the full-height slab of the triangle exchange, or for the two half-buffers of the owner-rows
            # layout (lower slab + finished rows, rows_alloc rows each)
            self.rows_alloc, _ = D.gram_block_rows(n_items, world, rank)
            n_rows = max(n_items, 2 * self.rows_alloc)
            _lib.check(lib.rt_ipc_alloc(4 * n_rows * D.slab_ld(n_items), C.byref(own), handle), "rt_ipc_alloc")
            self.own = int(own.value)
            payload = bytes(handle)
        except Exception as e:  # noqa: BLE001
            self.error = str(e)
        handles = [None] * world
        dist.all_gather_object(handles, payload, group=group)
        if self.error is None and any(h is None for h in handles):
            self.error = "a peer rank could not export its slab buffer"
        if self.error is None:
            try:
                for r in range(world):
                    if r == rank:
                        self.ptrs[r] = self.own
                        continue
                    p = C.c_void_p(0)
                    hb = (C.c_uint8 * 64).from_buffer_copy(handles[r])
                    _lib.check(lib.rt_ipc_open(hb, C.byref(p)), "rt_ipc_open")
                    self.ptrs[r] = int(p.value)
            except Exception as e:  # noqa: BLE001
                self.error = str(e)
        ok = t.tensor([0 if self.error else 1], dtype=t.int32, device=D.dev())
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.error = self.error or "a peer rank could not map the slab buffers"
            self.close()

    @classmethod
    def get(cls, n_items: int, rank: int, world: int, group=None) -> Optional["PeerSlabs"]:
        """The node's slab buffers for this shape, or None when CUDA IPC is not usable here (agreed by all
        ranks; the caller then exchanges the slabs with NCCL)."""
        key = (id(group), n_items, rank, world)
        if key not in cls._cache:
            for k in [k for k in cls._cache if k[0] == key[0]]:
                old = cls._cache.pop(k)
                if old is not None:
                    old.close()
            obj = cls(n_items, rank, world, group)
            cls._cache[key] = None if obj.error else obj
            if obj.error and rank == 0:
                import logging
                logging.warning(f"peer-memory slab exchange unavailable ({obj.error}); using NCCL broadcasts")
        return cls._cache[key]

    def rows_ptrs(self) -> list:
        """Addresses of every part's finished-rows buffer (second half of the IPC allocation)."""
        off = 4 * self.rows_alloc * D.slab_ld(self.n_items)
        return [int(p) + off for p in self.ptrs]

    def barrier(self) -> None:
        """Stream-ordered node barrier: the one-element all-reduce completes only after every rank has
        reached it on its stream."""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    def close(self) -> None:
        lib = _lib.load()
        D.torch().cuda.synchronize()
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                lib.rt_ipc_close(C.c_void_p(p))
        if self.own:
            lib.rt_ipc_free(C.c_void_p(self.own))
        self.ptrs, self.own = [], 0


def gram_sharded(X: D.DeviceMatrix, *, rank: int, world: int, group=None, exchange: str = "p2p", marks=None, live_cfg=None):
    """Item-item Gram matrix on every rank of the node: each rank computes its row slab of the rank-space
    lower triangle; ``exchange="p2p"`` pulls the other slabs over NVLink inside the mirror kernel
    (rt_gram_finish_p2p), ``exchange="nccl"`` broadcasts the slabs and mirrors locally."""
    mark = marks if marks is not None else (lambda name: None)
    if world <= 1:
        L = D.gram_lower(X)
        mark("gram_lower")
        G = D.gram_finish(L, live_cfg=live_cfg)    # (single GPU: see D.gram_finish for what live_cfg promises)
        mark("gram_finish")
        return G
    slabs = PeerSlabs.get(X.n_items, rank, world, group) if exchange == "p2p" else None
    if slabs is not None:
        slabs.barrier()   # nobody is still reading the slab of the previous fit
        L = D.gram_lower(X, part=rank, n_parts=world, raw_ptr=slabs.own)
        mark("gram_lower")
        slabs.barrier()   # every slab is complete
        G = D.gram_finish_p2p(L, slabs.ptrs, rank, X.n_items, barrier=slabs.barrier)
        mark("gram_finish_p2p")
        return G
    L = D.gram_lower(X, part=rank, n_parts=world)
    mark("gram_lower")
    exchange_slabs(L.Gp, L.cuts, group=group)
    mark("gram_exchange")
    G = D.gram_finish(L)
    mark("gram_finish")
    return G


def fit_owner_rows(X: D.DeviceMatrix, cfg: FitConfig, *, rank: int, world: int, group=None, want_sel: bool = False,
                   marks=None, slabs: Optional[PeerSlabs] = None, targets=None):
    """Multi-GPU fit without a full Gram exchange: every rank computes the lower-triangle part of the Gram rows it
    owns (block-cyclic in popularity-rank space), completes them by pulling the transposed columns below them out of
    the peers' slabs over NVLink (``rt_gram_pull_cols``: (N-1)/N^2 of the matrix per GPU), and solves its own targets;
    the few Gram entries between candidates that live in other ranks' rows are gathered from peer memory inside the
    solver.  Returns this rank's ``SolveResult`` (targets = the items of its blocks, restricted to ``targets`` -- a
    device int32 list, the same on every rank -- when given) or ``None`` when CUDA IPC is not usable on this node
    (agreed by every rank; the caller falls back to ``fit_sharded``)."""
    mark = marks if marks is not None else (lambda name: None)
    slabs = slabs if slabs is not None else PeerSlabs.get(X.n_items, rank, world, group)
    if slabs is None:
        return None
    I = X.n_items
    slabs.barrier()   # nobody is still reading this rank's buffers (pulls / solver gathers of the previous fit)
    rank_of, orig_of = D.gram_lower_blocks(X, rank, world, slabs.own)
    mark("gram_lower")
    slabs.barrier()   # every lower slab is complete
    D.gram_pull_cols(slabs.ptrs, rank, I)
    rows = slabs.rows_ptrs()
    D.gram_unpermute_rows(slabs.own, rows[rank], slabs.rows_alloc, I, rank_of)
    G = D.GramRows(rows, D.gram_row_slots(rank_of, world), D.slab_ld(I))
    tg = D.block_targets(orig_of, I, rank, world)
    if targets is not None:
        t = D.torch()
        tg = tg[t.isin(tg, targets)].contiguous()
    mark("gram_rows")
    slabs.barrier()   # every rank's rows are complete
    res = D.solve(G, I, tg, cfg, want_sel=want_sel)
    mark("solve")
    return res


def fit_sharded(X: D.DeviceMatrix, cfg: FitConfig, *, rank: int = 0, world: int = 1, group=None,
                targets=None, want_sel: bool = False, strided: bool = False, exchange: str = "p2p"):
    """Gram row slab of this rank -> slab exchange fused with the mirror (P2P) -> unpermute -> solve this rank's targets.

    Returns ``(res, (j0, j1))`` where ``res`` holds this rank's columns of W (``SolveResult``).
    With ``world == 1`` this is the single-GPU bulk fit.
    """
    t = D.require_cuda()
    I = X.n_items
    j0, j1 = item_shard(I, rank, world)
    # K3: this rank's row slab of the rank-space lower triangle (rows balanced by multiply-add count),
    # exchanged over peer memory inside the mirror kernel
    G = gram_sharded(X, rank=rank, world=world, group=group, exchange=exchange)
    if targets is not None:
        tg = targets
    elif strided and world > 1:
        tg = item_stride(I, rank, world)
    else:
        tg = t.arange(j0, j1, dtype=t.int32, device=D.dev())
    res = D.solve(G, I, tg, cfg, want_sel=want_sel)
    del G
    return res, (j0, j1)


def all_gather_ragged(x, counts: list, group=None):
    """All-gather of 1-D tensors of unequal length: rank r contributes ``x[:counts[r]]`` (``counts`` is known
    to every rank); returns the list of per-rank pieces.  One padded ``all_gather_into_tensor``."""
    import torch.distributed as dist
    t = D.torch()
    world = len(counts)
    rank = dist.get_rank(group)
    n_max = max(max(counts), 1)
    buf = t.zeros(n_max, dtype=x.dtype, device=x.device)
    if counts[rank]:
        buf[:counts[rank]] = x[:counts[rank]]
    out = t.empty((world, n_max), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out.view(-1), buf, group=group)
    return [out[r, :counts[r]] for r in range(world)]


def gather_solve_results(res: D.SolveResult, world: int, group=None) -> D.SolveResult:
    """All-gather (NCCL) of the per-rank solver outputs: every rank ends up with the (target, rows, values)
    triples of all targets, ready for ``w_merge``.  Two collectives: the (n_targets, n_pairs) header, then ONE packed
    int32 buffer per rank -- targets | cnt | stats | off (as two int32 halves) | rows | value bits -- instead of one
    all-gather per array (seven launches and seven padded staging buffers at ~3 MB of payload, SCALE_r01)."""
    if world <= 1:
        return res
    import torch.distributed as dist
    t = D.torch()
    dev = res.targets.device
    T = int(res.targets.numel())
    P_ = int(res.n_pairs)
    meta = t.tensor([T, P_], dtype=t.int64, device=dev)
    metas = t.empty((world, 2), dtype=t.int64, device=dev)
    dist.all_gather_into_tensor(metas.view(-1), meta, group=group)
    metas = metas.cpu().tolist()
    Ts = [int(m[0]) for m in metas]
    Ps = [int(m[1]) for m in metas]
    Tm, Pm = max(max(Ts), 1), max(max(Ps), 1)
    width = 8 * Tm + 2 * Pm               # targets, cnt: Tm each; stats: 4 Tm; off: 2 Tm; rows, vals: Pm each
    buf = t.zeros(width, dtype=t.int32, device=dev)
    if T:
        buf[0:T] = res.targets
        buf[Tm:Tm + T] = res.cnt[:T]
        buf[2 * Tm:2 * Tm + 4 * T] = res.stats.reshape(-1)[:4 * T]
        buf[6 * Tm:6 * Tm + 2 * T] = res.off[:T].view(t.int32)
    if P_:
        buf[8 * Tm:8 * Tm + P_] = res.rows[:P_]
        buf[8 * Tm + Pm:8 * Tm + Pm + P_] = res.vals[:P_].view(t.int32)
    out = t.empty((world, width), dtype=t.int32, device=dev)
    dist.all_gather_into_tensor(out.view(-1), buf, group=group)
    g_t, g_cnt, g_stats, g_off, g_rows, g_vals = [], [], [], [], [], []
    base = 0
    for r in range(world):
        o, Tr, Pr = out[r], Ts[r], Ps[r]
        g_t.append(o[0:Tr]); g_cnt.append(o[Tm:Tm + Tr]); g_stats.append(o[2 * Tm:2 * Tm + 4 * Tr])
        g_off.append(o[6 * Tm:6 * Tm + 2 * Tr].view(t.int64) + base)
        g_rows.append(o[8 * Tm:8 * Tm + Pr]); g_vals.append(o[8 * Tm + Pm:8 * Tm + Pm + Pr].view(t.float32))
        base += Pr
    return D.SolveResult(t.cat(g_t), t.cat(g_off), t.cat(g_cnt), t.cat(g_rows), t.cat(g_vals), None,
                         t.cat(g_stats).view(-1, 4), res.rows_sorted, base)


def recommend_sharded(X: D.DeviceMatrix, users, W_shard: D.DeviceW, j_range: Tuple[int, int], k: int,
                      filter_interacted: bool, mode: int, *, world: int = 1, group=None):
    """Item-partitioned scoring: K6 on this rank's item columns for every user, all-gather of the per-rank
    (ids, scores) lists, K7 merge.  ``W_shard`` needs only the columns ``j_range`` of W."""
    t = D.require_cuda()
    ids, scores, cnt = D.recommend(X, users, W_shard, k, filter_interacted, mode, j_range[0], j_range[1])
    if world == 1:
        return ids, scores, cnt
    import torch.distributed as dist
    Q = int(users.numel())
    all_ids = t.empty((world, Q, k), dtype=t.int32, device=D.dev())
    all_sc = t.empty((world, Q, k), dtype=t.float32, device=D.dev())
    dist.all_gather_into_tensor(all_ids.view(-1), ids.contiguous().view(-1), group=group)
    dist.all_gather_into_tensor(all_sc.view(-1), scores.contiguous().view(-1), group=group)
    return D.topk_merge(all_ids, all_sc, world, Q, k)


def recommend_query_sharded(X: D.DeviceMatrix, users, W: D.DeviceW, k: int, filter_interacted: bool, mode: int, *,
                            rank: int = 0, world: int = 1, group=None):
    """Query-partitioned scoring: every rank holds all of W (a few MB), scores its slice of the query list
    against every item column (no merge step), and the finished top-k lists are all-gathered so that every
    rank returns the lists of all users."""
    t = D.require_cuda()
    if world == 1:
        return D.recommend(X, users, W, k, filter_interacted, mode)
    import torch.distributed as dist
    Q = int(users.numel())
    cuts = query_cuts(X.rptr, users, world)
    bounds = [(cuts[r], cuts[r + 1]) for r in range(world)]
    q0, q1 = bounds[rank]
    ids, scores, cnt = D.recommend(X, users[q0:q1], W, k, filter_interacted, mode)
    m = max(b - a for a, b in bounds)
    pack = t.zeros((max(m, 1), 2 * k + 1), dtype=t.int32, device=D.dev())
    n = q1 - q0
    if n:
        pack[:n, :k] = ids
        pack[:n, k:2 * k] = scores.view(t.int32)
        pack[:n, 2 * k] = cnt
    out = t.empty((world, max(m, 1), 2 * k + 1), dtype=t.int32, device=D.dev())
    dist.all_gather_into_tensor(out.view(-1), pack.view(-1), group=group)
    parts = [out[r, :b - a] for r, (a, b) in enumerate(bounds)]
    full = t.cat(parts, 0)
    return full[:, :k].contiguous(), full[:, k:2 * k].contiguous().view(t.float32), full[:, 2 * k].contiguous()
