"""Array-level pipeline over the C-ABI: events -> store -> X -> W -> top-k, all device resident.

This is the layer ``UserItemInteractions`` / ``SLIM`` sit on, exposed so that callers who already
hold event columns on the device (bench.py, the multi-GPU driver) can run the same kernels
without the host object model.  Sharding helpers for ``torch.distributed`` live here too: item
columns are partitioned across ranks, X is replicated, Gram rows are all-gathered (NCCL) before
the solves, and per-rank top-k lists are all-gathered and merged by ``rt_topk_merge``.
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


# ------------------------------------------------------------------------------------------ sharding
def item_shard(n_items: int, rank: int, world: int, cptr_host: Optional[np.ndarray] = None) -> Tuple[int, int]:
    """Contiguous item range [j0, j1) owned by ``rank``.  With the CSC column pointer given, the cut
    points balance the Gram work (entries per column weighted by mean row length is close to
    balancing stored entries); otherwise equal item counts."""
    if world <= 1:
        return 0, n_items
    if cptr_host is None:
        cuts = [(n_items * r) // world for r in range(world + 1)]
    else:
        nnz = int(cptr_host[-1])
        cuts = [0]
        for r in range(1, world):
            cuts.append(int(np.searchsorted(cptr_host, (nnz * r) // world, side="left")))
        cuts.append(n_items)
        cuts = [min(max(c, 0), n_items) for c in cuts]
        for r in range(1, world + 1):
            cuts[r] = max(cuts[r], cuts[r - 1])
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


def fit_sharded(X: D.DeviceMatrix, cfg: FitConfig, *, rank: int = 0, world: int = 1, group=None,
                targets=None, want_sel: bool = False):
    """Gram row slab of this rank -> all-gather of the slabs (NCCL) -> mirror/unpermute -> solve this rank's targets.

    Returns ``(res, (j0, j1))`` where ``res`` holds this rank's columns of W (``SolveResult``).
    With ``world == 1`` this is the single-GPU bulk fit.
    """
    t = D.require_cuda()
    I = X.n_items
    cptr_host = X.cptr.cpu().numpy() if world > 1 else None
    j0, j1 = item_shard(I, rank, world, cptr_host)
    # K3: this rank's row slab of the rank-space lower triangle (rows balanced by multiply-add count)
    L = D.gram_lower(X, part=rank, n_parts=world)
    if world > 1:
        exchange_slabs(L.Gp, L.cuts, group=group)
    G = D.gram_finish(L)
    del L
    if targets is None:
        tg = t.arange(j0, j1, dtype=t.int32, device=D.dev())
    else:
        tg = targets
    res = D.solve(G, I, tg, cfg, want_sel=want_sel)
    del G
    return res, (j0, j1)


def recommend_sharded(X: D.DeviceMatrix, users, W_shard: D.DeviceW, j_range: Tuple[int, int], k: int,
                      filter_interacted: bool, mode: int, *, world: int = 1, group=None):
    """K6 on this rank's item columns, all-gather of the per-rank (ids, scores) lists, K7 merge."""
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
