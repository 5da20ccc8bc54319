"""Device-resident containers and thin Python wrappers over the C-ABI (``include/rtrec_b200.h``).

PyTorch is used here only as plumbing: device allocations (``torch.empty``), the current CUDA
stream and host<->device copies.  Every computation is a call into ``librtrec_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import FitConfig, RtrecB200Error, check

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RtrecB200Error("no CUDA device visible: rtrec_b200 runs its hot path on the GPU only (no CPU fallback)")
    _lib.load()
    return t


def dev():
    t = require_cuda()
    return t.device("cuda", t.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch().cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def to_dev(a: np.ndarray, dtype=None):
    t = require_cuda()
    x = t.from_numpy(np.ascontiguousarray(a if dtype is None else a.astype(dtype, copy=False)))
    return x.to(dev(), non_blocking=False)


def empty(n, dtype):
    t = require_cuda()
    return t.empty(int(n), dtype=dtype, device=dev())


def zeros(n, dtype):
    t = require_cuda()
    return t.zeros(int(n), dtype=dtype, device=dev())


# --------------------------------------------------------------------------------------------
@dataclass
class DeviceMatrix:
    """Interaction matrix X on the device in CSR and CSC (plus the COO column ids of the CSC order).
    float32 values / int32 indices like the reference's scipy matrices (interactions.py:276,303)."""
    n_users: int
    n_items: int
    nnz: int
    rptr: object
    ridx: object
    rval: object
    cptr: object
    cidx: object
    cval: object
    ccol: object
    nonneg: bool

    @property
    def shape(self) -> Tuple[int, int]:
        return self.n_users, self.n_items

    @staticmethod
    def from_scipy(X) -> "DeviceMatrix":
        """Upload a host scipy matrix (operator-level API used by tests / HybridSlimFM-style callers)."""
        import scipy.sparse as sp
        t = require_cuda()
        csr = sp.csr_matrix(X, dtype=np.float32)
        csr.sum_duplicates()        # (the kernels rely on one stored entry per (user, item), as the device store guarantees)
        csr.sort_indices()
        csc = sp.csc_matrix(csr)
        csc.sort_indices()
        n_users, n_items = csr.shape
        ccol = np.repeat(np.arange(n_items, dtype=np.int32), np.diff(csc.indptr).astype(np.int64))
        nonneg = bool(csr.nnz == 0 or csr.data.min() >= 0)
        return DeviceMatrix(
            n_users, n_items, int(csr.nnz),
            to_dev(csr.indptr, np.int32), to_dev(csr.indices, np.int32), to_dev(csr.data, np.float32),
            to_dev(csc.indptr, np.int32), to_dev(csc.indices, np.int32), to_dev(csc.data, np.float32),
            to_dev(ccol, np.int32), nonneg)

    def to_scipy_csr(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.rval[:self.nnz].cpu().numpy(), self.ridx[:self.nnz].cpu().numpy(),
                              self.rptr.cpu().numpy()), shape=self.shape)

    def to_scipy_csc(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.cval[:self.nnz].cpu().numpy(), self.cidx[:self.nnz].cpu().numpy(),
                              self.cptr.cpu().numpy()), shape=self.shape)


@dataclass
class DeviceW:
    """Item-similarity matrix W (n_items x n_items) on the device: CSC (by target column, what the
    reference exposes as ``item_similarity``) and CSR (by source item, what scoring streams)."""
    n_items: int
    nnz: int
    wptr: object
    widx: object
    wval: object
    wrptr: object
    wridx: object
    wrval: object
    packs: Optional[dict] = None   # (j_begin, j_end) -> ScorePack, built on first use

    def to_scipy_csc(self, dtype=np.float32):
        import scipy.sparse as sp
        return sp.csc_matrix((self.wval[:self.nnz].cpu().numpy().astype(dtype), self.widx[:self.nnz].cpu().numpy(),
                              self.wptr.cpu().numpy()), shape=(self.n_items, self.n_items))

    @staticmethod
    def from_scipy(W) -> "DeviceW":
        import scipy.sparse as sp
        t = require_cuda()
        csc = sp.csc_matrix(W, dtype=np.float32)
        csc.sort_indices()
        n = csc.shape[1]
        wptr, widx, wval = to_dev(csc.indptr, np.int32), to_dev(csc.indices, np.int32), to_dev(csc.data, np.float32)
        return _finish_w(n, int(csc.nnz), wptr, widx, wval)


def _finish_w(n_items: int, nnz: int, wptr, widx, wval) -> DeviceW:
    t = torch()
    wrptr = empty(n_items + 1, t.int32)
    wridx = empty(max(nnz, 1), t.int32)
    wrval = empty(max(nnz, 1), t.float32)
    check(_lib.load().rt_transpose(n_items, n_items, ptr(wptr), ptr(widx), ptr(wval), nnz, ptr(wrptr), ptr(wridx),
                                   ptr(wrval), stream_ptr()), "rt_transpose")
    return DeviceW(n_items, nnz, wptr, widx, wval, wrptr, wridx, wrval)


# --------------------------------------------------------------------------------------------
_rng_cache: Dict[Tuple[int, int], object] = {}


def rng_table(seed: int, n: int):
    """xorshift32 draw table (device); cached per (device, seed) and grown on demand."""
    t = require_cuda()
    key = (t.cuda.current_device(), int(seed))
    cur = _rng_cache.get(key)
    if cur is not None and cur.numel() >= n:
        return cur
    n_alloc = int(n * 1.25) + 1024
    out = empty(n_alloc, t.int32)
    check(_lib.load().rt_rng_table(C.c_uint32(seed), n_alloc, ptr(out), stream_ptr()), "rt_rng_table")
    _rng_cache[key] = out
    return out


def gram(X: DeviceMatrix, e_begin: int = 0, e_end: Optional[int] = None, out=None):
    """Dense item-item Gram matrix G (n_items x n_items float32) on the device (K3)."""
    t = require_cuda()
    I = X.n_items
    if out is None:
        out = t.zeros((I, I), dtype=t.float32, device=dev())
    e_end = X.nnz if e_end is None else e_end
    check(_lib.load().rt_gram_rows(ptr(X.ccol), ptr(X.cidx), ptr(X.cval), int(e_begin), int(e_end), ptr(X.rptr),
                                   ptr(X.ridx), ptr(X.rval), ptr(out), I, stream_ptr()), "rt_gram_rows")
    return out


@dataclass
class GramLower:
    """Lower triangle of the Gram matrix in popularity-rank space (output of rt_gram_lower)."""
    Gp: object        # float32 [I, I] device, rows [cuts[part], cuts[part+1]) filled
    rank_of: object   # int32 [I]
    orig_of: object   # int32 [I]
    cuts: list        # row cut points, len n_parts + 1


def slab_ld(n_items: int) -> int:
    """Leading dimension of a raw slab buffer: padded so that rows start 16-byte aligned (wide P2P loads)."""
    return (int(n_items) + 3) // 4 * 4


def gram_lower(X: DeviceMatrix, part: int = 0, n_parts: int = 1, raw_ptr: Optional[int] = None) -> GramLower:
    """Row slab ``part`` of the rank-space lower triangle.  ``raw_ptr``: write into this raw device buffer
    (I rows of ``slab_ld(I)`` floats, e.g. a CUDA IPC buffer shared with the peer ranks) instead of a fresh
    torch tensor; the returned ``Gp`` is then that integer address."""
    t = require_cuda()
    lib = _lib.load()
    I = X.n_items
    if raw_ptr is None:
        Gp = t.zeros((I, I), dtype=t.float32, device=dev())
        gp_ptr, ld = ptr(Gp), I
    else:
        Gp = int(raw_ptr)
        gp_ptr, ld = C.c_void_p(Gp), slab_ld(I)
        check(lib.rt_memset(gp_ptr, 0, 4 * I * ld, stream_ptr()), "rt_memset")
    rank_of = empty(I, t.int32)
    orig_of = empty(I, t.int32)
    cuts = (C.c_int32 * (n_parts + 1))()
    check(lib.rt_gram_lower(X.n_users, I, ptr(X.cptr), ptr(X.cidx), ptr(X.cval), ptr(X.rptr), ptr(X.ridx),
                            ptr(X.rval), X.nnz, int(part), int(n_parts), gp_ptr, ld, ptr(rank_of), ptr(orig_of),
                            cuts, stream_ptr()), "rt_gram_lower")
    return GramLower(Gp, rank_of, orig_of, list(cuts))


def gram_finish_p2p(L: GramLower, slab_ptrs: Sequence[int], part: int, n_items: int, out=None, barrier=None):
    """Fused slab exchange + mirror over peer memory in two balanced phases, then the un-permuted symmetric G
    (rt_gram_finish_p2p phases 0, 1, 2).  ``slab_ptrs[p]``: address of rank p's slab buffer in this process
    (own buffer or IPC mapping).  ``barrier``: stream-ordered node barrier, called between the two pull
    phases (omit when all slabs live in this process, as in the single-GPU test)."""
    t = require_cuda()
    lib = _lib.load()
    n_parts = len(slab_ptrs)
    if out is None:
        out = t.empty((n_items, n_items), dtype=t.float32, device=dev())
    arr = (C.c_void_p * n_parts)(*[C.c_void_p(int(p)) for p in slab_ptrs])
    cuts = (C.c_int32 * (n_parts + 1))(*[int(c) for c in L.cuts])
    for phase in (0, 1, 2):
        check(lib.rt_gram_finish_p2p(n_items, arr, n_parts, int(part), cuts, slab_ld(n_items), ptr(L.rank_of),
                                     ptr(L.orig_of), ptr(out), n_items, phase, stream_ptr()), "rt_gram_finish_p2p")
        if phase == 0 and barrier is not None:
            barrier()
    return out


def gram_block_rows(n_items: int, n_parts: int, part: int = 0) -> Tuple[int, int]:
    """(rows_alloc, rows_own) of the block-cyclic owner-rows layout (rt_gram_block_rows)."""
    ra, ro = C.c_int32(0), C.c_int32(0)
    check(_lib.load().rt_gram_block_rows(int(n_items), int(n_parts), int(part), C.byref(ra), C.byref(ro)), "rt_gram_block_rows")
    return int(ra.value), int(ro.value)


def gram_lower_blocks(X: DeviceMatrix, part: int, n_parts: int, slab_ptr: int):
    """Lower-triangle part of the rows ``part`` owns (block-cyclic, rank space) into the raw buffer ``slab_ptr``
    ([rows_alloc, slab_ld(I)] floats, zero-filled here).  Returns (rank_of, orig_of)."""
    t = require_cuda()
    lib = _lib.load()
    I = X.n_items
    ld = slab_ld(I)
    rows_alloc, _ = gram_block_rows(I, n_parts, part)
    check(lib.rt_memset(C.c_void_p(int(slab_ptr)), 0, 4 * rows_alloc * ld, stream_ptr()), "rt_memset")
    rank_of = empty(I, t.int32)
    orig_of = empty(I, t.int32)
    check(lib.rt_gram_lower_blocks(X.n_users, I, ptr(X.cptr), ptr(X.cidx), ptr(X.cval), ptr(X.rptr), ptr(X.ridx),
                                   ptr(X.rval), X.nnz, int(part), int(n_parts), C.c_void_p(int(slab_ptr)), ld,
                                   ptr(rank_of), ptr(orig_of), stream_ptr()), "rt_gram_lower_blocks")
    return rank_of, orig_of


def gram_pull_cols(slab_ptrs: Sequence[int], part: int, n_items: int) -> None:
    """Completes the own rows: transposed pull of the triangle columns below them out of every part's slab."""
    arr = (C.c_void_p * len(slab_ptrs))(*[C.c_void_p(int(p)) for p in slab_ptrs])
    check(_lib.load().rt_gram_pull_cols(int(n_items), arr, len(slab_ptrs), int(part), slab_ld(n_items), stream_ptr()),
          "rt_gram_pull_cols")


def gram_unpermute_rows(slab_ptr: int, rows_ptr: int, n_rows: int, n_items: int, rank_of) -> None:
    """Columns of the own rows back to item ids: rows[r][i] = slab[r][rank_of[i]]."""
    ld = slab_ld(n_items)
    check(_lib.load().rt_gram_unpermute_rows(int(n_rows), int(n_items), C.c_void_p(int(slab_ptr)), ld, ptr(rank_of),
                                             C.c_void_p(int(rows_ptr)), ld, stream_ptr()), "rt_gram_unpermute_rows")


def gram_row_slots(rank_of, n_parts: int):
    t = require_cuda()
    I = int(rank_of.numel())
    slots = empty(I, t.int32)
    check(_lib.load().rt_gram_row_slots(I, ptr(rank_of), int(n_parts), ptr(slots), stream_ptr()), "rt_gram_row_slots")
    return slots


def block_targets(orig_of, n_items: int, part: int, n_parts: int):
    """Item ids of the targets ``part`` owns under the block-cyclic layout (device int32), in local-row order."""
    t = require_cuda()
    _, rows_own = gram_block_rows(n_items, n_parts, part)
    l = t.arange(rows_own, dtype=t.int64, device=dev())
    jp = ((l >> 6) * n_parts + part) * 64 + (l & 63)   # only the globally last block can be partial, and it is last here too
    return orig_of.long()[jp].to(t.int32)


def gram_finish(L: GramLower, out=None, live_cfg: Optional[FitConfig] = None):
    """Mirror + back to item ids.  ``live_cfg``: the matrix is for ONE ``solve`` call with this configuration and without
    candidate lists in or out; rows the solver provably never reads (bulk fit with feature selection: targets without a
    live coordinate) then get only their diagonal entry (rt_gram_finish_live) and ``out._rt_live_only`` is set."""
    t = require_cuda()
    I = L.Gp.shape[0]
    if out is None:
        out = t.empty((I, I), dtype=t.float32, device=dev())
    # the rows pass through shared memory on their way back to item ids: their largest off-diagonal entry comes for free
    # and rides along on the tensor (``solve`` hands it to the solver, which then skips trivial targets without a read)
    rowmax = empty(I, t.float32)
    has = C.c_int32(0)
    live = C.c_int32(0)
    if live_cfg is not None:
        check(_lib.load().rt_gram_finish_live(I, ptr(L.Gp), I, ptr(L.rank_of), ptr(L.orig_of), ptr(out), I, ptr(rowmax),
                                              C.byref(live_cfg), C.byref(has), C.byref(live), stream_ptr()), "rt_gram_finish_live")
    else:
        check(_lib.load().rt_gram_finish_rowmax(I, ptr(L.Gp), I, ptr(L.rank_of), ptr(L.orig_of), ptr(out), I, ptr(rowmax),
                                                C.byref(has), stream_ptr()), "rt_gram_finish_rowmax")
    out._rt_rowmax = rowmax if has.value else None
    out._rt_live_only = bool(live.value)
    return out


def gram_full(X: DeviceMatrix, out=None, live_cfg: Optional[FitConfig] = None):
    """Dense symmetric item-item Gram matrix G in item ids (K3, third generation)."""
    return gram_finish(gram_lower(X), out=out, live_cfg=live_cfg)


def gram_cols(X: DeviceMatrix, j_begin: int = 0, j_end: Optional[int] = None, out=None):
    """Gram rows of the target columns [j_begin, j_end) with shared-memory accumulators (rt_gram)."""
    t = require_cuda()
    I = X.n_items
    if out is None:
        out = t.zeros((I, I), dtype=t.float32, device=dev())
    j_end = I if j_end is None else j_end
    check(_lib.load().rt_gram(X.n_users, I, ptr(X.cptr), ptr(X.cidx), ptr(X.cval), ptr(X.rptr), ptr(X.ridx), ptr(X.rval),
                              int(j_begin), int(j_end), ptr(out), I, stream_ptr()), "rt_gram")
    return out


def set_option(name: str, value: int) -> None:
    """Tuning switches.  ``score_impl``: 3 = packed scoring kernel (default), 2 / 1 = earlier generations."""
    global _score_impl, _score_tc
    if name == "score_tc":
        _score_tc = int(value)
        return
    if name == "fit_pruned":
        global _fit_pruned
        _fit_pruned = int(value)
        return
    if name == "score_impl":
        _score_impl = int(value)
        if int(value) == 3:
            return
    check(_lib.load().rt_set_option(name.encode(), int(value)), "rt_set_option")


@dataclass
class SolveResult:
    targets: object   # int32 device [T]
    off: object       # int64 device [T]
    cnt: object       # int32 device [T]
    rows: object      # int32 device
    vals: object      # float32 device
    sel: Optional[object]    # int32 device [T, nn] or None
    stats: object     # int32 device [T, 4]
    rows_sorted: bool
    n_pairs: int


@dataclass
class GramRows:
    """Gram matrix in the owner-rows layout of the multi-GPU fit (rt_gram_lower_blocks ... rt_gram_row_slots):
    ``bases[p]`` = address of part p's row buffer in this process (own memory or CUDA IPC mapping), ``slots`` =
    int32 [I] device, (owner << 24) | local row, ``ld`` = common leading dimension."""
    bases: Sequence[int]
    slots: object
    ld: int


def solve(G, n_items: int, targets, cfg: FitConfig, sel_in=None, want_sel: bool = False) -> SolveResult:
    """Batched ElasticNet solves on the Gram matrix (K4).  ``G``: dense [I, ld] tensor or ``GramRows``."""
    t = require_cuda()
    lib = _lib.load()
    T = int(targets.numel())
    nn = int(cfg.nn)
    NU = min(nn, n_items) if nn > 0 else n_items
    rng = rng_table(int(cfg.seed), int(cfg.max_iter) * NU + 64)
    off = empty(max(T, 1), t.int64)
    cnt = zeros(max(T, 1), t.int32)
    stats = zeros(max(T, 1) * 4, t.int32)
    sel_out = empty(max(T * nn, 1), t.int32) if (nn > 0 and want_sel) else None
    cap = T * NU if nn > 0 else max(T * min(NU, 256), 1024)
    needed = C.c_int64(0)
    while True:
        rows = empty(max(cap, 1), t.int32)
        vals = empty(max(cap, 1), t.float32)
        if isinstance(G, GramRows):
            arr = (C.c_void_p * len(G.bases))(*[C.c_void_p(int(b)) for b in G.bases])
            rc = lib.rt_slim_solve_rows(arr, len(G.bases), ptr(G.slots), int(G.ld), n_items, ptr(targets), T, C.byref(cfg),
                                        ptr(sel_in), ptr(rng), rng.numel(), ptr(sel_out), ptr(off), ptr(cnt), ptr(rows),
                                        ptr(vals), cap, C.byref(needed), ptr(stats), stream_ptr())
        else:
            if getattr(G, "_rt_live_only", False) and (sel_in is not None or want_sel or not (cfg.nn > 0 and cfg.skip_trivial)):
                raise ValueError("this Gram matrix holds only the rows of a bulk fit with feature selection (gram_finish live_cfg)")
            rowmax = getattr(G, "_rt_rowmax", None)
            if rowmax is not None and not cfg.rowmax_ptr:
                cfg = FitConfig.from_buffer_copy(cfg)
                cfg.rowmax_ptr = rowmax.data_ptr()
            rc = lib.rt_slim_solve(ptr(G), G.stride(0), n_items, ptr(targets), T, C.byref(cfg), ptr(sel_in), ptr(rng),
                                   rng.numel(), ptr(sel_out), ptr(off), ptr(cnt), ptr(rows), ptr(vals), cap, C.byref(needed),
                                   ptr(stats), stream_ptr())
        if rc == _lib.RT_ERR_CAPACITY and needed.value > cap:
            cap = int(needed.value)
            continue
        check(rc, "rt_slim_solve")
        break
    return SolveResult(targets, off, cnt, rows, vals, sel_out.view(T, nn) if sel_out is not None else None,
                       stats.view(-1, 4)[:T], nn == 0, int(needed.value))


_fit_pruned = 1     # 0 = never use the pruned all-features fit (set_option("fit_pruned", 0))
last_pruned_rows = None


def fit_pruned(X: DeviceMatrix, targets, cfg: FitConfig) -> Optional[SolveResult]:
    """Fit without the dense Gram matrix (rt_slim_fit_pruned): only the Gram rows of the items that can take part in a
    non-zero solution are formed.  Applies to all-features fits and to bulk fits with feature selection
    (``cfg.skip_trivial``).  Returns the SolveResult of ``targets``, or None when the path does not apply (negative data,
    non-positive coefficients, a merge into an existing W with feature selection, or too many candidate rows): the
    caller then builds G."""
    global last_pruned_rows
    if not _fit_pruned or not cfg.positive or not cfg.nonneg or X.nnz == 0 or (int(cfg.nn) != 0 and not cfg.skip_trivial):
        return None
    t = require_cuda()
    lib = _lib.load()
    T = int(targets.numel())
    if T == 0:
        return None
    I = X.n_items
    nn = int(cfg.nn)
    NU = min(nn, I) if nn > 0 else I
    rng = rng_table(int(cfg.seed), int(cfg.max_iter) * NU + 64)
    off = empty(T, t.int64); cnt = zeros(T, t.int32); stats = zeros(T * 4, t.int32)
    cap = T * NU if nn > 0 else max(T * min(I, 256), 1024)
    needed, used, n_rows = C.c_int64(0), C.c_int32(0), C.c_int32(0)
    while True:
        rows = empty(cap, t.int32); vals = empty(cap, t.float32)
        rc = lib.rt_slim_fit_pruned(X.n_users, I, ptr(X.cptr), ptr(X.cidx), ptr(X.cval), ptr(X.ccol), ptr(X.rptr), ptr(X.ridx),
                                    ptr(X.rval), X.nnz, ptr(targets), T, C.byref(cfg), ptr(rng), rng.numel(), ptr(off), ptr(cnt),
                                    ptr(rows), ptr(vals), cap, C.byref(needed), ptr(stats), C.byref(used), C.byref(n_rows),
                                    stream_ptr())
        if rc == _lib.RT_ERR_CAPACITY and needed.value > cap:
            cap = int(needed.value)
            continue
        check(rc, "rt_slim_fit_pruned")
        break
    last_pruned_rows = int(n_rows.value)
    if not used.value:
        return None
    return SolveResult(targets, off, cnt, rows, vals, None, stats.view(-1, 4), nn == 0, int(needed.value))


def w_merge(old: Optional[DeviceW], n_items: int, res: SolveResult) -> DeviceW:
    """Assemble / merge W with the reference's LIL-assignment semantics (K5)."""
    t = require_cuda()
    lib = _lib.load()
    T = int(res.targets.numel())
    old_nnz = old.nnz if old is not None else 0
    cap = old_nnz + res.n_pairs + 16
    wptr = empty(n_items + 1, t.int32)
    widx = empty(cap, t.int32)
    wval = empty(cap, t.float32)
    nnz = C.c_int64(0)
    check(lib.rt_w_merge(n_items, ptr(old.wptr) if old else None, ptr(old.widx) if old else None,
                         ptr(old.wval) if old else None, old.n_items if old else 0, ptr(res.targets), T, ptr(res.off),
                         ptr(res.cnt), ptr(res.rows), ptr(res.vals), 1 if res.rows_sorted else 0, ptr(wptr), ptr(widx),
                         ptr(wval), cap, C.byref(nnz), stream_ptr()), "rt_w_merge")
    n = int(nnz.value)
    return _finish_w(n_items, n, wptr, widx[:max(n, 1)], wval[:max(n, 1)])


@dataclass
class ScorePack:
    """Bank-striped ELL pack of the heavy rows of W for one item range (score3.cu)."""
    heavy_of: object   # int32 [n_items]
    ell_off: object    # int32 [n_heavy * n_tiles + 1]
    ell: object        # int32 [n_groups * 64] or None
    n_heavy: int
    n_groups: int
    tile: int
    n_tiles: int


PACK_MIN_ROW = 768  # W rows shorter than this stay on the CSR path (one or two 512-thread steps)
_score_impl = 3     # 3 = packed kernel (default), 2 / 1 = earlier generations through rt_slim_recommend


def score_pack(W: DeviceW, j_begin: int, j_end: int) -> ScorePack:
    if W.packs is None:
        W.packs = {}
    key = (int(j_begin), int(j_end))
    hit = W.packs.get(key)
    if hit is not None:
        return hit
    t = require_cuda()
    lib = _lib.load()
    I = W.n_items
    heavy_of = empty(I, t.int32)
    heavy_list = empty(I, t.int32)
    n_heavy, tile, n_tiles, n_groups = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int64(0)
    check(lib.rt_score_tile(I, key[0], key[1], C.byref(tile), C.byref(n_tiles)), "rt_score_tile")
    ell_off = empty(I * max(int(n_tiles.value), 1) + 1, t.int32)
    check(lib.rt_w_pack_plan(ptr(W.wrptr), ptr(W.wridx), I, key[0], key[1], PACK_MIN_ROW, ptr(heavy_of), ptr(heavy_list),
                             ptr(ell_off), C.byref(n_heavy), C.byref(tile), C.byref(n_tiles), C.byref(n_groups),
                             stream_ptr()), "rt_w_pack_plan")
    ell = None
    if n_heavy.value > 0 and n_groups.value > 0:
        ell = empty(int(n_groups.value) * 64, t.int32)
        check(lib.rt_w_pack_fill(ptr(W.wrptr), ptr(W.wridx), ptr(W.wrval), I, key[0], key[1], ptr(heavy_list),
                                 int(n_heavy.value), ptr(ell_off), ptr(ell), int(n_groups.value), stream_ptr()),
              "rt_w_pack_fill")
    pack = ScorePack(heavy_of, ell_off[:int(n_heavy.value) * int(n_tiles.value) + 1], ell, int(n_heavy.value),
                     int(n_groups.value), int(tile.value), int(n_tiles.value))
    W.packs[key] = pack
    return pack


@dataclass
class TcPack:
    """Tensor-core form of the heavy rows of W (score_tc.cu): three K-major bf16 planes + a dense fp32 copy."""
    bt: object        # uint8 buffer [3 * i_pad * 64 * 2]
    wd: object        # float32 [n_heavy, n_items]
    heavy_of: object  # int32 [n_items]
    n_heavy: int
    w_nonneg: bool


TC_MIN_QUERIES = 4096   # below this the exact kernel is used (a launch of the tensor-core pipeline does not pay)
TC_KMAX = 16
_score_tc = 1           # 0 = never use the tensor-core scoring path (set_option("score_tc", 0))


def tc_pack(W: DeviceW) -> Optional[TcPack]:
    """Tensor-core pack of W, or None when W does not qualify (no heavy row, more than 64 heavy rows, negative weights)."""
    if W.packs is None:
        W.packs = {}
    if "tc" in W.packs:
        return W.packs["tc"]
    t = require_cuda()
    lib = _lib.load()
    pk = score_pack(W, 0, W.n_items)
    out = None
    if 0 < pk.n_heavy <= 64:
        i_pad, bt_b, wd_b = C.c_int32(0), C.c_int64(0), C.c_int64(0)
        check(lib.rt_tc_pack_size(W.n_items, pk.n_heavy, C.byref(i_pad), C.byref(bt_b), C.byref(wd_b)), "rt_tc_pack_size")
        bt = t.empty(int(bt_b.value), dtype=t.uint8, device=dev())
        wd = t.empty(int(wd_b.value) // 4, dtype=t.float32, device=dev())
        heavy_list = t.nonzero(pk.heavy_of >= 0).flatten().to(t.int32)   # ascending item id = heavy slot order (pack_index_kernel)
        nonneg = C.c_int32(0)
        check(lib.rt_tc_pack_build(ptr(W.wrptr), ptr(W.wridx), ptr(W.wrval), W.nnz, W.n_items, ptr(heavy_list), pk.n_heavy,
                                   ptr(bt), ptr(wd), C.byref(nonneg), stream_ptr()), "rt_tc_pack_build")
        out = TcPack(bt, wd, pk.heavy_of, pk.n_heavy, bool(nonneg.value))
    W.packs["tc"] = out
    return out


def values_bf16_exact(X: DeviceMatrix) -> Tuple[bool, bool]:
    """(all stored values of X >= 0, all exactly representable in bf16); cached on the matrix."""
    hit = getattr(X, "_bf16_info", None)
    if hit is None:
        nn_, ex = C.c_int32(0), C.c_int32(0)
        check(_lib.load().rt_values_bf16_exact(ptr(X.rval), X.nnz, C.byref(nn_), C.byref(ex), stream_ptr()), "rt_values_bf16_exact")
        hit = (bool(nn_.value), bool(ex.value))
        X._bf16_info = hit
    return hit


TC_REDO_MIN_SLOTS, TC_REDO_DIV = 256, 16   # no_sync scoring: max(256, Q // 16) slots for users handed back by the tensor-core path


def recommend_tc(X: DeviceMatrix, users, W: DeviceW, k: int, filter_interacted: bool, mode: int, debug_scores: bool = False,
                 no_sync: bool = False):
    """Scoring with the heavy rows of W on the tensor cores (score_tc.cu); scores agree with the exact kernels to ~1e-6
    relative.  Returns device (ids, scores, cnt) like ``recommend``, or None when the preconditions do not hold.  Users the
    fast path cannot finish (light-row table overflow, dense-mode lists with fewer than k positive scores) are re-scored
    by the exact kernel.

    ``no_sync``: the host does not wait for the list of those users.  A fixed number of slots (``Q // 16``, at least 256)
    is re-scored -- the flagged users first, the spare slots into a scratch row -- and a fourth result, the device count
    of flagged users, tells the caller afterwards whether the slots were enough (``> slots``: call again without
    ``no_sync``).  Lets a caller queue many batches back to back (``SLIMElastic.recommend_lists``)."""
    t = require_cuda()
    Q = int(users.numel())
    if k > TC_KMAX or Q == 0:
        return None
    pk = tc_pack(W)
    if pk is None or not pk.w_nonneg:
        return None
    x_nonneg, x_exact = values_bf16_exact(X)
    if not x_nonneg:
        return None
    ids = empty((Q + 1) * k, t.int32); scores = empty((Q + 1) * k, t.float32); cnt = empty(Q + 1, t.int32)   # (+1: scratch row)
    tc_ids = empty(Q * 32, t.int32); tc_sc = empty(Q * 32, t.float32); tc_cnt = empty(Q, t.int32)
    fb = empty(Q, t.int32)
    i_pad = (W.n_items + 127) // 128 * 128
    dbg = t.zeros((Q, i_pad), dtype=t.float32, device=dev()) if debug_scores else None
    check(_lib.load().rt_slim_recommend_tc(ptr(X.rptr), ptr(X.ridx), ptr(X.rval), ptr(users), Q, ptr(W.wrptr), ptr(W.wridx),
                                           ptr(W.wrval), ptr(pk.heavy_of), pk.n_heavy, ptr(pk.bt), ptr(pk.wd), W.n_items, int(k),
                                           1 if filter_interacted else 0, int(mode), 1 if x_exact else 3, ptr(tc_ids), ptr(tc_sc),
                                           ptr(tc_cnt), ptr(ids), ptr(scores), ptr(cnt), ptr(fb), ptr(dbg), stream_ptr()),
          "rt_slim_recommend_tc")
    ids, scores = ids.view(Q + 1, k), scores.view(Q + 1, k)
    global _score_tc
    if no_sync and not debug_scores:
        slots = min(Q, max(TC_REDO_MIN_SLOTS, Q // TC_REDO_DIV))
        flag = (fb != 0).to(t.int32)
        hit, redo = t.topk(flag, slots)
        redo = t.where(hit != 0, redo, Q)                       # spare slots: user id -1 (no query), result to the scratch row
        keep, _score_tc = _score_tc, 0
        try:
            r_ids, r_sc, r_cnt = recommend(X, t.nn.functional.pad(users, (0, 1), value=-1)[redo].contiguous(), W, k,
                                           filter_interacted, mode)
        finally:
            _score_tc = keep
        ids[redo] = r_ids; scores[redo] = r_sc; cnt[redo] = r_cnt
        return ids[:Q], scores[:Q], cnt[:Q], (flag.sum(dtype=t.int32), slots)
    redo = t.nonzero(fb).flatten()
    if int(redo.numel()):
        keep, _score_tc = _score_tc, 0
        try:
            r_ids, r_sc, r_cnt = recommend(X, users[redo].contiguous(), W, k, filter_interacted, mode)
        finally:
            _score_tc = keep
        ids[redo] = r_ids; scores[redo] = r_sc; cnt[redo] = r_cnt
    ids, scores, cnt = ids[:Q], scores[:Q], cnt[:Q]
    if debug_scores:
        return ids, scores, cnt, dbg, (tc_ids.view(Q, 32), tc_sc.view(Q, 32), tc_cnt), redo
    return ids, scores, cnt


def recommend(X: DeviceMatrix, users, W: DeviceW, k: int, filter_interacted: bool, mode: int,
              j_begin: int = 0, j_end: Optional[int] = None, no_sync: bool = False):
    """Fused scoring + filter + top-k for a batch of users (K6).  Returns device (ids, scores, cnt).  Large batches over the
    whole item range go to the tensor-core path when W and X qualify (``recommend_tc``), everything else to the exact kernels.
    ``no_sync``: see ``recommend_tc``; a fourth result (None, or (device count, slots)) is returned."""
    t = require_cuda()
    Q = int(users.numel())
    if (_score_tc and _score_impl == 3 and Q >= TC_MIN_QUERIES and k <= TC_KMAX and j_begin == 0
            and (j_end is None or j_end == W.n_items)):
        out = recommend_tc(X, users, W, k, filter_interacted, mode, no_sync=no_sync)
        if out is not None:
            return out
    if no_sync:
        return recommend(X, users, W, k, filter_interacted, mode, j_begin, j_end) + (None,)
    ids = empty(max(Q * k, 1), t.int32)
    scores = empty(max(Q * k, 1), t.float32)
    cnt = empty(max(Q, 1), t.int32)
    j_end = W.n_items if j_end is None else j_end
    if _score_impl == 3 and j_end > j_begin and Q > 0:
        pk = score_pack(W, j_begin, j_end)
        check(_lib.load().rt_slim_recommend_packed(ptr(X.rptr), ptr(X.ridx), ptr(X.rval), ptr(users), Q, ptr(W.wrptr),
                                                   ptr(W.wridx), ptr(W.wrval), ptr(pk.heavy_of), ptr(pk.ell_off), ptr(pk.ell),
                                                   W.n_items, int(j_begin), int(j_end), int(k),
                                                   1 if filter_interacted else 0, int(mode), ptr(ids), ptr(scores),
                                                   ptr(cnt), stream_ptr()), "rt_slim_recommend_packed")
    else:
        check(_lib.load().rt_slim_recommend(ptr(X.rptr), ptr(X.ridx), ptr(X.rval), ptr(users), Q, ptr(W.wrptr), ptr(W.wridx),
                                            ptr(W.wrval), W.n_items, int(j_begin), int(j_end), int(k),
                                            1 if filter_interacted else 0, int(mode), ptr(ids), ptr(scores), ptr(cnt),
                                            stream_ptr()), "rt_slim_recommend")
    return ids.view(Q, k) if Q else ids[:0].view(0, k), scores.view(Q, k) if Q else scores[:0].view(0, k), cnt[:Q]


def recommend_candidates(X: DeviceMatrix, users, W: DeviceW, cand, k: int):
    t = require_cuda()
    Q = int(users.numel())
    pos = empty(max(Q * k, 1), t.int32)
    scores = empty(max(Q * k, 1), t.float32)
    cnt = empty(max(Q, 1), t.int32)
    check(_lib.load().rt_slim_recommend_candidates(ptr(X.rptr), ptr(X.ridx), ptr(X.rval), ptr(users), Q, ptr(W.wptr),
                                                   ptr(W.widx), ptr(W.wval), W.n_items, ptr(cand), int(cand.numel()),
                                                   int(k), ptr(pos), ptr(scores), ptr(cnt), stream_ptr()),
          "rt_slim_recommend_candidates")
    return pos.view(Q, k), scores.view(Q, k), cnt[:Q]


def similar(W: DeviceW, items, k: int):
    t = require_cuda()
    Q = int(items.numel())
    ids = empty(max(Q * k, 1), t.int32)
    scores = empty(max(Q * k, 1), t.float32)
    cnt = empty(max(Q, 1), t.int32)
    check(_lib.load().rt_slim_similar(ptr(W.wptr), ptr(W.widx), ptr(W.wval), W.n_items, ptr(items), Q, int(k), ptr(ids),
                                      ptr(scores), ptr(cnt), stream_ptr()), "rt_slim_similar")
    return ids.view(Q, k), scores.view(Q, k), cnt[:Q]


def topk_merge(ids, scores, n_shards: int, n_query: int, k: int):
    t = require_cuda()
    out_ids = empty(max(n_query * k, 1), t.int32)
    out_scores = empty(max(n_query * k, 1), t.float32)
    cnt = empty(max(n_query, 1), t.int32)
    check(_lib.load().rt_topk_merge(ptr(ids), ptr(scores), n_shards, n_query, k, ptr(out_ids), ptr(out_scores), ptr(cnt),
                                    stream_ptr()), "rt_topk_merge")
    return out_ids.view(n_query, k), out_scores.view(n_query, k), cnt[:n_query]
