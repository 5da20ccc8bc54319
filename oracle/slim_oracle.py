"""CPU oracle for rtrec's SLIM hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package ``rtrec_b200`` never
does and has no CPU fallback.

It restates, on the CPU, what the reference computes on this path (citations are
``/root/reference`` paths, rtrec 0.2.7):

* ``UserItemInteractions`` ingest / decay / clip / matrix export
  (rtrec/utils/interactions.py:15-119, 259-303)  -> :class:`StoreOracle`, :func:`fold_events`
* ``SLIMElastic.fit`` / ``partial_fit_items`` / ``FeatureSelectionWrapper.fit``
  (rtrec/models/internal/slim_elastic.py:139-154, 229-281, 510-564) -> :class:`SlimOracle`
  with the ElasticNet solve in ``slim_oracle.c`` (a port of scikit-learn 1.9.0
  ``sparse_enet_coordinate_descent``; sklearn is a third-party dependency of the reference that
  is not vendored in its tree -- see the C file header).
* scoring + top-k (slim_elastic.py:628-818) and ``similar_items`` (:820-857).

Pinned by tests/test_oracle_pin.py against (a) the installed scikit-learn and (b) golden
vectors produced by importing the real reference in the build container
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libslim_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile slim_oracle.c + gram_model.c with gcc (see oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("slim_oracle.c", "gram_model.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(_LIB_PATH) < os.path.getmtime(f) for f in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.so_enet_solve.restype = ctypes.c_int
        _lib.so_fit_columns.restype = ctypes.c_int
        _lib.so_feature_scores.restype = ctypes.c_int
        _lib.so_num_threads.restype = ctypes.c_int
        _lib.gm_gram.restype = ctypes.c_int
        _lib.gm_fit_columns.restype = ctypes.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def sklearn_seed(random_state: int = 43) -> int:
    """Seed the Cython solver draws per fit: ``rng.randint(0, RAND_R_MAX)`` on a fresh
    ``RandomState(random_state)`` (sklearn/linear_model/_cd_fast.pyx:748)."""
    return int(np.random.RandomState(random_state).randint(0, 2**31 - 1))


def enet_solve(X: sp.csc_matrix, y: np.ndarray, alpha=0.1, l1_ratio=0.1, max_iter=100, tol=1e-4,
               random_state=43, positive=True) -> Tuple[np.ndarray, int, float]:
    """One ``ElasticNet(fit_intercept=False, selection='random').fit(X, y)`` on sparse X."""
    X = sp.csc_matrix(X, dtype=np.float32)
    X.sort_indices()
    y = np.ascontiguousarray(y, dtype=np.float32)
    w = np.zeros(X.shape[1], dtype=np.float32)
    gap = ctypes.c_float(0)
    data = np.ascontiguousarray(X.data, dtype=np.float32)
    idx = np.ascontiguousarray(X.indices, dtype=np.int32)
    ptr = np.ascontiguousarray(X.indptr, dtype=np.int32)
    it = lib().so_enet_solve(ctypes.c_int(X.shape[0]), ctypes.c_int(X.shape[1]), _p(data), _p(idx), _p(ptr),
                             _p(y), ctypes.c_double(alpha), ctypes.c_double(l1_ratio), ctypes.c_int(max_iter),
                             ctypes.c_double(tol), ctypes.c_uint32(sklearn_seed(random_state)),
                             ctypes.c_int(int(positive)), _p(w), ctypes.byref(gap))
    return w, int(it), float(gap.value)


def _canon(X, fmt: str):
    M = sp.csc_matrix(X, dtype=np.float32) if fmt == "csc" else sp.csr_matrix(X, dtype=np.float32)
    M.sort_indices()
    return (np.ascontiguousarray(M.data, dtype=np.float32), np.ascontiguousarray(M.indices, dtype=np.int32),
            np.ascontiguousarray(M.indptr, dtype=np.int32))


def feature_scores(X_csc: sp.csc_matrix, j: int) -> np.ndarray:
    """``X.T.dot(y)`` with column j zeroed, same fp32 accumulation order as scipy (slim_elastic.py:141)."""
    cd, ci, cp = _canon(X_csc, "csc")
    X_csr = sp.csc_matrix((cd, ci, cp), shape=X_csc.shape).tocsr()
    rd, ri, rp = _canon(X_csr, "csr")
    s = np.zeros(X_csc.shape[1], dtype=np.float32)
    lib().so_feature_scores(ctypes.c_int(X_csc.shape[0]), ctypes.c_int(X_csc.shape[1]), _p(cd), _p(ci), _p(cp),
                            _p(rd), _p(ri), _p(rp), ctypes.c_int(int(j)), _p(s))
    return s


def fit_columns(X_csc, targets: Sequence[int], nn: Optional[int] = None, sel_in: Optional[np.ndarray] = None,
                alpha=0.1, l1_ratio=0.1, max_iter=100, tol=1e-4, random_state=43, positive=True,
                n_threads: int = 1):
    """Per-column body of SLIMElastic.fit for ``targets``.

    Returns ``(cols, sel, stats)``: ``cols[t] = (rows int32, vals float32)`` exactly what the
    reference iterates at slim_elastic.py:273/556 (FS path: all ``nn`` picks incl. zeros, in
    pick order; otherwise the non-zeros in ascending row), ``sel`` the candidate order used
    (``None`` without FS), ``stats[t] = (n_iter, visits, gap_evals, nnz_selected)``.
    """
    n_users, n_items = X_csc.shape
    cd, ci, cp = _canon(X_csc, "csc")
    csr = sp.csc_matrix((cd, ci, cp), shape=(n_users, n_items)).tocsr()
    rd, ri, rp = _canon(csr, "csr")
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    T = len(targets)
    nn_c = int(nn) if nn else 0
    cap = nn_c if nn_c > 0 else n_items
    out_rows = np.zeros((T, cap), dtype=np.int32)
    out_vals = np.zeros((T, cap), dtype=np.float32)
    out_cnt = np.zeros(T, dtype=np.int32)
    stats = np.zeros((T, 4), dtype=np.int64)
    sel_out = np.full((T, nn_c), -1, dtype=np.int32) if nn_c > 0 else None
    if sel_in is not None:
        sel_in = np.ascontiguousarray(sel_in, dtype=np.int32)
        assert sel_in.shape == (T, nn_c)
    rc = lib().so_fit_columns(
        ctypes.c_int(n_users), ctypes.c_int(n_items), _p(cd), _p(ci), _p(cp), _p(rd), _p(ri), _p(rp),
        ctypes.c_int(T), _p(targets), ctypes.c_int(nn_c), _p(sel_in),
        ctypes.c_double(alpha), ctypes.c_double(l1_ratio), ctypes.c_int(max_iter), ctypes.c_double(tol),
        ctypes.c_uint32(sklearn_seed(random_state)), ctypes.c_int(int(positive)), ctypes.c_int(n_threads),
        _p(sel_out), ctypes.c_int(cap), _p(out_rows), _p(out_vals), _p(out_cnt), _p(stats))
    if rc != 0:
        raise RuntimeError(f"so_fit_columns failed rc={rc}")
    cols = [(out_rows[t, :out_cnt[t]].copy(), out_vals[t, :out_cnt[t]].copy()) for t in range(T)]
    return cols, sel_out, stats


def gram_model_gram(X_csc) -> np.ndarray:
    """Dense fp32 Gram matrix exactly as gram_model.c accumulates it (CPU model of the device path)."""
    n_users, n_items = X_csc.shape
    cd, ci, cp = _canon(X_csc, "csc")
    rd, ri, rp = _canon(sp.csc_matrix((cd, ci, cp), shape=(n_users, n_items)).tocsr(), "csr")
    G = np.zeros((n_items, n_items), dtype=np.float32)
    lib().gm_gram(ctypes.c_int(n_users), ctypes.c_int(n_items), _p(cd), _p(ci), _p(cp), _p(rd), _p(ri), _p(rp), _p(G))
    return G


def gram_model_fit_columns(X_csc, targets, nn=None, sel_in=None, alpha=0.1, l1_ratio=0.1, max_iter=100, tol=1e-4,
                           random_state=43, positive=True, G=None):
    """CPU model of the device Gram-form replay (oracle/gram_model.c).  Same return shape as
    :func:`fit_columns`; ``stats[t] = (n_iter, draws, gap_evals, live_size)``."""
    n_users, n_items = X_csc.shape
    if G is None:
        G = gram_model_gram(X_csc)
    nonneg = bool(X_csc.nnz == 0 or X_csc.data.min() >= 0)
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    T = len(targets)
    nn_c = int(nn) if nn else 0
    cap = nn_c if nn_c > 0 else n_items
    out_rows = np.zeros((T, cap), dtype=np.int32)
    out_vals = np.zeros((T, cap), dtype=np.float32)
    out_cnt = np.zeros(T, dtype=np.int32)
    stats = np.zeros((T, 4), dtype=np.int64)
    sel_out = np.full((T, nn_c), -1, dtype=np.int32) if nn_c > 0 else None
    if sel_in is not None:
        sel_in = np.ascontiguousarray(sel_in, dtype=np.int32)
    rc = lib().gm_fit_columns(
        ctypes.c_int(n_users), ctypes.c_int(n_items), _p(G), ctypes.c_int(T), _p(targets), ctypes.c_int(nn_c),
        _p(sel_in), ctypes.c_double(alpha), ctypes.c_double(l1_ratio), ctypes.c_int(max_iter), ctypes.c_double(tol),
        ctypes.c_uint32(sklearn_seed(random_state)), ctypes.c_int(int(positive)), ctypes.c_int(int(nonneg)),
        _p(sel_out), ctypes.c_int(cap), _p(out_rows), _p(out_vals), _p(out_cnt), _p(stats))
    if rc != 0:
        raise RuntimeError(f"gm_fit_columns failed rc={rc}")
    cols = [(out_rows[t, :out_cnt[t]].copy(), out_vals[t, :out_cnt[t]].copy()) for t in range(T)]
    return cols, sel_out, stats


def numpy_candidates(X_csc, j: int, nn: int) -> np.ndarray:
    """The reference's own pick ``np.argsort(s)[-1:-1-n:-1]`` (numpy's unstable tie order)."""
    s = feature_scores(X_csc, j)
    return np.argsort(s)[-1:-1 - nn:-1].astype(np.int32)


# --------------------------------------------------------------------------- W bookkeeping
class SlimOracle:
    """``SLIMElastic`` restated (serial path).  ``item_similarity`` is a scipy CSC like the
    reference's attribute; values float32 (the reference's float64 container after a serial
    ``fit`` holds the same float32 numbers, slim_elastic.py:252)."""

    def __init__(self, config: dict | None = None):
        config = config or {}
        self.alpha = config.get("alpha", 0.1)
        self.l1_ratio = config.get("l1_ratio", 0.1)
        self.positive_only = config.get("positive_only", True)
        self.max_iter = config.get("max_iter", 100)
        self.tol = config.get("tol", 1e-4)
        self.random_state = config.get("random_state", 43)
        self.nn_feature_selection = config.get("nn_feature_selection", None)
        self.item_similarity: Optional[sp.csc_matrix] = None
        self.n_threads = config.get("n_threads", 1)
        self.last_sel = None
        self.last_stats = None

    def _solve(self, X, items, sel_in=None):
        return fit_columns(X, items, self.nn_feature_selection, sel_in, self.alpha, self.l1_ratio, self.max_iter,
                           self.tol, self.random_state, self.positive_only, self.n_threads)

    @staticmethod
    def _apply(cols_old: Dict[int, Dict[int, float]], j: int, rows, vals):
        # LIL ``M[i, j] = v`` semantics: non-zero inserts/overwrites, zero deletes (slim_elastic.py:273-274)
        col = cols_old.setdefault(int(j), {})
        for i, v in zip(rows.tolist(), vals.tolist()):
            if v != 0.0:
                col[int(i)] = v
            else:
                col.pop(int(i), None)

    @staticmethod
    def _to_csc(cols: Dict[int, Dict[int, float]], n_items: int) -> sp.csc_matrix:
        indptr = np.zeros(n_items + 1, dtype=np.int64)
        idx, dat = [], []
        for j in range(n_items):
            c = cols.get(j)
            if c:
                ks = sorted(c)
                idx.extend(ks)
                dat.extend(c[k] for k in ks)
            indptr[j + 1] = len(idx)
        return sp.csc_matrix((np.asarray(dat, dtype=np.float32), np.asarray(idx, dtype=np.int32),
                              indptr.astype(np.int32)), shape=(n_items, n_items))

    @staticmethod
    def _from_csc(W: Optional[sp.csc_matrix]) -> Dict[int, Dict[int, float]]:
        cols: Dict[int, Dict[int, float]] = {}
        if W is None:
            return cols
        W = W.tocsc()
        for j in range(W.shape[1]):
            a, b = W.indptr[j], W.indptr[j + 1]
            if b > a:
                cols[j] = dict(zip(W.indices[a:b].tolist(), W.data[a:b].astype(np.float32).tolist()))
        return cols

    def fit(self, X, sel_in=None) -> "SlimOracle":
        n_items = X.shape[1]
        items = np.arange(n_items, dtype=np.int32)
        res, self.last_sel, self.last_stats = self._solve(X, items, sel_in)
        cols: Dict[int, Dict[int, float]] = {}
        for j, (rows, vals) in zip(items, res):
            self._apply(cols, j, rows, vals)
        self.item_similarity = self._to_csc(cols, n_items)
        return self

    def partial_fit_items(self, X, updated_items: Sequence[int], sel_in=None) -> "SlimOracle":
        n_items = X.shape[1]
        cols = self._from_csc(self.item_similarity)  # tolil().resize(): old entries kept
        items = np.asarray(list(updated_items), dtype=np.int32)
        res, self.last_sel, self.last_stats = self._solve(X, items, sel_in)
        for j, (rows, vals) in zip(items, res):
            self._apply(cols, j, rows, vals)
        self.item_similarity = self._to_csc(cols, n_items)
        return self

    # -- scoring (slim_elastic.py:628-818) --
    def scores_dense(self, user_ids: Sequence[int], X_csr: sp.csr_matrix) -> np.ndarray:
        if self.item_similarity is None:
            raise RuntimeError("Model must be fitted before calling predict.")
        return np.asarray((X_csr[list(user_ids), :] @ self.item_similarity.astype(np.float32)).todense(),
                          dtype=np.float32)

    def recommend_batch(self, user_ids, X_csr, candidate_item_ids=None, top_k=10, filter_interacted=True,
                        dense_output=True, ret_scores=False):
        if self.item_similarity is None:
            raise RuntimeError("Model must be fitted before calling batch_recommend.")
        W = self.item_similarity.astype(np.float32)
        out = []
        if candidate_item_ids is not None:
            S = np.asarray((X_csr[list(user_ids), :] @ W[:, candidate_item_ids]).todense(), dtype=np.float32)
            for r in range(len(user_ids)):
                order = np.argsort(S[r], kind="stable")[-top_k:][::-1]
                items = [candidate_item_ids[i] for i in order]
                out.append((items, S[r][order]) if ret_scores else items)
            return out
        for u in user_ids:
            row = X_csr[u, :]
            if dense_output:
                s = np.asarray((row @ W).todense(), dtype=np.float32).ravel()
                if filter_interacted:
                    s[row.indices] = -np.inf
                top = np.argsort(s, kind="stable")[-top_k:][::-1]
                top = top[s[top] != -np.inf] if len(top) else top
                out.append((top.tolist(), s[top]) if ret_scores else top.tolist())
            else:
                sc = sp.csr_matrix(row @ W)
                pairs = list(zip(sc.indices.tolist(), sc.data.tolist()))
                if filter_interacted:
                    seen = set(row.indices.tolist())
                    pairs = [(i, v) for i, v in pairs if i not in seen]
                pairs = sorted(pairs, key=lambda x: x[1], reverse=True)[:top_k]
                if ret_scores:
                    out.append(([i for i, _ in pairs], np.array([v for _, v in pairs], dtype=np.float32)))
                else:
                    out.append([i for i, _ in pairs])
        return out

    def similar_items(self, item_id: int, top_k=10):
        if self.item_similarity is None:
            raise RuntimeError("Model must be fitted before calling similar_items.")
        col = self.item_similarity[:, item_id]
        idx, val = col.indices, col.data
        m = idx != item_id
        idx, val = idx[m], val[m]
        order = np.argsort(-val, kind="stable")[:top_k]
        return list(zip(idx[order].tolist(), val[order].tolist()))


# --------------------------------------------------------------------------- interaction store
class StoreOracle:
    """Event-at-a-time restatement of ``UserItemInteractions`` (interactions.py:15-119,259-303).
    Pure Python: small inputs only.  ``fold_events`` below is the vectorised equivalent."""

    def __init__(self, min_value=-5, max_value=10, decay_in_days=None):
        assert max_value > min_value
        self.min_value, self.max_value = min_value, max_value
        self.decay_rate = None if decay_in_days is None else 1.0 - (math.log(2) / decay_in_days)
        self.pairs: Dict[Tuple[int, int], Tuple[float, float]] = {}
        self.max_user_id = 0
        self.max_item_id = 0
        self.max_timestamp = 0.0

    def _decay(self, value, ts):
        if self.decay_rate is None:
            return value
        return value * self.decay_rate ** ((self.max_timestamp - ts) / 86400.0)

    def rating(self, u, i):
        cur = self.pairs.get((u, i))
        if cur is None or cur[0] == 0.0:
            return 0.0
        return self._decay(cur[0], cur[1])

    def add(self, u, i, ts, delta=1.0, upsert=False):
        self.max_timestamp = max(self.max_timestamp, ts + 1.0)
        if upsert:
            self.pairs[(u, i)] = (delta, ts)
        else:
            new = self.rating(u, i) + delta
            new = max(self.min_value, min(new, self.max_value))
            self.pairs[(u, i)] = (new, ts)
        self.max_user_id = max(self.max_user_id, u)
        self.max_item_id = max(self.max_item_id, i)

    def to_coo_arrays(self, select_items=None, select_users=None):
        rows, cols, data = [], [], []
        for (u, i), (v, ts) in self.pairs.items():
            if select_items is not None and i not in select_items:
                continue
            if select_users and u not in select_users:
                continue
            rows.append(u); cols.append(i); data.append(self._decay(v, ts))
        return rows, cols, data

    def to_csc(self, select_items=None):
        r, c, d = self.to_coo_arrays(select_items=select_items)
        return sp.csc_matrix((d, (r, c)), shape=(self.max_user_id + 1, self.max_item_id + 1), dtype="float32")

    def to_csr(self, select_users=None):
        r, c, d = self.to_coo_arrays(select_users=select_users)
        return sp.csr_matrix((d, (r, c)), shape=(self.max_user_id + 1, self.max_item_id + 1), dtype="float32")


def fold_events(users, items, ts, delta, *, upsert=False, min_value=-5, max_value=10, decay_in_days=None,
                state=None):
    """Vectorised restatement of a batch of ``add_interaction`` calls in arrival order (SURVEY.md
    Appendix B).  ``state`` = ``(keys int64 sorted, values f64, stamps f64, max_ts, max_u, max_i)`` from a
    previous call or ``None``.  Returns the new state in the same form.
    """
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    ts = np.asarray(ts, dtype=np.float64)
    delta = np.asarray(delta, dtype=np.float64)
    rate = None if decay_in_days is None else 1.0 - (math.log(2) / decay_in_days)
    if state is None:
        keys0 = np.zeros(0, np.int64); v0 = np.zeros(0); s0 = np.zeros(0); mts = 0.0; mu = 0; mi = 0
    else:
        keys0, v0, s0, mts, mu, mi = state
    n = len(users)
    if n == 0:
        return keys0, v0, s0, mts, mu, mi
    T = np.maximum.accumulate(np.maximum(ts + 1.0, mts))  # max_timestamp seen by event k (interactions.py:99)
    key = (users << 32) | items
    order = np.argsort(key, kind="stable")
    ks, tss, ds, Ts = key[order], ts[order], delta[order], T[order]
    first = np.ones(n, dtype=bool); first[1:] = ks[1:] != ks[:-1]
    grp = np.cumsum(first) - 1
    ukeys = ks[first]
    # previous state per group
    pos = np.searchsorted(keys0, ukeys)
    pos_c = np.minimum(pos, max(len(keys0) - 1, 0))
    found = (pos < len(keys0)) & (keys0[pos_c] == ukeys) if len(keys0) else np.zeros(len(ukeys), bool)
    cur_v = np.where(found, v0[pos_c] if len(keys0) else 0.0, 0.0)
    cur_s = np.where(found, s0[pos_c] if len(keys0) else 0.0, 0.0)
    rank = np.arange(n) - np.flatnonzero(first)[grp]
    for r in range(int(rank.max()) + 1):
        m = rank == r
        g = grp[m]
        if upsert:
            cur_v[g] = ds[m]
        else:
            prev = cur_v[g]
            if rate is not None:
                dec = prev * rate ** ((Ts[m] - cur_s[g]) / 86400.0)
                prev = np.where(prev == 0.0, 0.0, dec)
            cur_v[g] = np.maximum(min_value, np.minimum(prev + ds[m], max_value))
        cur_s[g] = tss[m]
    # merge
    keep_old = np.ones(len(keys0), dtype=bool)
    if len(keys0):
        keep_old[pos_c[found]] = False
    keys = np.concatenate([keys0[keep_old], ukeys])
    vals = np.concatenate([v0[keep_old], cur_v])
    stamps = np.concatenate([s0[keep_old], cur_s])
    o = np.argsort(keys, kind="stable")
    return (keys[o], vals[o], stamps[o], float(T[-1]), int(max(mu, users.max())), int(max(mi, items.max())))


def state_to_matrix(state, *, decay_in_days=None, fmt="csc", select_items=None):
    keys, vals, stamps, mts, mu, mi = state
    rate = None if decay_in_days is None else 1.0 - (math.log(2) / decay_in_days)
    u = (keys >> 32).astype(np.int64)
    i = (keys & 0xFFFFFFFF).astype(np.int64)
    x = vals if rate is None else vals * rate ** ((mts - stamps) / 86400.0)
    if select_items is not None:
        m = np.isin(i, np.asarray(list(select_items), dtype=np.int64))
        u, i, x = u[m], i[m], x[m]
    shape = (mu + 1, mi + 1)
    M = sp.coo_matrix((x.astype(np.float32), (u, i)), shape=shape)
    return M.tocsc() if fmt == "csc" else M.tocsr()
