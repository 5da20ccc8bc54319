"""``SLIMElastic`` -- the SLIM operator, B200-native.

Same constructor config, methods and error behaviour as the reference operator
(/root/reference/rtrec/models/internal/slim_elastic.py:156-857); the arithmetic runs in the
hand-written kernels of ``librtrec_b200.so``:

    fit / fit_in_parallel / partial_fit_items  ->  rt_gram_lower/finish (K3) + rt_slim_solve (K4) + rt_w_merge (K5)
    recommend / recommend_batch                ->  rt_slim_recommend[_candidates] (K6)
    similar_items                              ->  rt_slim_similar (K8)

Matrices may be passed as scipy CSR/CSC (uploaded once per call, like the reference's operator
API) or as :class:`rtrec_b200.device.DeviceMatrix` (what ``SLIM`` passes: no host copy).
There is no CPU path: without a GPU these methods raise ``RtrecB200Error``.
"""
from __future__ import annotations

import gc
import logging
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp
from numpy import ndarray

from ... import _lib
from ... import device as D
from ..._lib import FitConfig, RT_TOPK_DENSE, RT_TOPK_SPARSE


_PINNED: Dict[str, Any] = {}  # pinned host staging buffers for results (grow-only)
_K_DEVICE = 128               # longest top-k list the fused scoring kernels keep per user (score3.cu KMAX3)


def sklearn_seed(random_state) -> int:
    """What sklearn's Cython solver seeds its xorshift32 with for ``random_state``
    (``check_random_state(rs).randint(0, RAND_R_MAX)``, _cd_fast.pyx:748)."""
    if random_state is None or isinstance(random_state, (int, np.integer)):
        rs = np.random.RandomState(None if random_state is None else int(random_state))
    else:
        rs = random_state
    return int(rs.randint(0, 2**31 - 1))


class SLIMElastic:
    def __init__(self, config: dict = {}):
        self.optim_name = config.get("optim", "cd")
        self.eta0 = config.get("eta0", 0.001)
        self.alpha = config.get("alpha", 0.1)
        self.l1_ratio = config.get("l1_ratio", 0.1)
        self.positive_only = config.get("positive_only", True)
        self.max_iter = config.get("max_iter", 100)
        self.tol = config.get("tol", 1e-4)
        self.random_state = config.get("random_state", 43)
        self.nn_feature_selection = config.get("nn_feature_selection", None)
        if self.nn_feature_selection is not None:
            assert int(self.nn_feature_selection) > 0, f"n_neighbors must be a positive integer: {self.nn_feature_selection}"
        # SPMD multi-GPU mode (not in the reference): with ``distributed=True`` and an initialised torch.distributed
        # process group (one process per GPU, every rank making the same calls with the same data), fits are sharded
        # by item column (pipeline.fit_owner_rows) and bulk scoring by query user; every rank ends up with the full W
        # and returns the full result, so the API reads exactly like the single-GPU one.
        self.distributed = bool(config.get("distributed", False))
        # how bulk scoring sees the query list in SPMD mode: "replicated" = every rank passes the same users, the list is
        # cut by stored entries, every rank scores its cut and all ranks return all lists (an all-gather); "local" = every
        # rank passes ITS OWN users (the caller shards the queries, e.g. users[rank::world]) and gets their lists back --
        # no collective and no foreign Python objects on any rank
        self.distributed_queries = str(config.get("distributed_queries", "replicated"))
        assert self.distributed_queries in ("replicated", "local"), self.distributed_queries
        self._W: Optional[D.DeviceW] = None
        self._W_host: Optional[sp.csc_matrix] = None  # lazy mirror / pending upload
        self.last_fit_stats: Optional[np.ndarray] = None
        self.last_fit_sel: Optional[np.ndarray] = None
        self.last_fit_targets: Optional[np.ndarray] = None
        self.keep_fit_details = bool(config.get("keep_fit_details", False))

    # ------------------------------------------------------------------ item_similarity mirror
    @property
    def item_similarity(self) -> Optional[sp.csc_matrix]:
        """scipy CSC view of W (the attribute the reference exposes, slim_elastic.py:193)."""
        if self._W_host is None and self._W is not None:
            self._W_host = self._W.to_scipy_csc()
        return self._W_host

    @item_similarity.setter
    def item_similarity(self, W: Optional[sp.spmatrix]) -> None:
        self._W = None
        self._W_host = None if W is None else sp.csc_matrix(W, dtype=np.float32)

    def _set_host_similarity(self, W: sp.spmatrix) -> None:
        """Take over a fitted W from the host (a model file written by the reference: scipy CSC, float64 after a
        serial ``fit`` -- the values were computed in float32 by the solver, so the cast is exact)."""
        self.item_similarity = W

    def _device_W(self) -> Optional[D.DeviceW]:
        if self._W is None and self._W_host is not None:
            self._W = D.DeviceW.from_scipy(self._W_host)
        return self._W

    # ------------------------------------------------------------------ helpers
    def _check_optim(self) -> None:
        if self.optim_name == "cd":
            return
        if self.optim_name == "sgd":
            raise NotImplementedError("optim='sgd' (SGDRegressor, slim_elastic.py:209-222) is not on the accelerated path")
        raise ValueError(f"Invalid Optimizer name: {self.optim_name}")

    def _config(self, X: D.DeviceMatrix, into_empty_w: bool = True) -> FitConfig:
        """``into_empty_w``: the result will be assembled into an empty W (bulk fit) -- columns that are zero before the first
        sweep may then come back without their nn zero coefficients (rt_fit_config.skip_trivial); a merge into an existing
        W needs them (a returned zero deletes a stale entry, slim_elastic.py:533-538)."""
        nn = int(self.nn_feature_selection) if self.nn_feature_selection is not None else 0
        return FitConfig(alpha=float(self.alpha), l1_ratio=float(self.l1_ratio), tol=float(self.tol),
                         max_iter=int(self.max_iter), positive=1 if self.positive_only else 0,
                         seed=sklearn_seed(self.random_state), nn=nn, n_samples=int(X.n_users),
                         nonneg=1 if X.nonneg else 0, skip_trivial=1 if into_empty_w else 0)

    @staticmethod
    def _as_device(interaction_matrix, allow=("csc", "csr"), err="Interaction matrix must be a scipy.sparse.csr_matrix or scipy.sparse.csc_matrix.") -> D.DeviceMatrix:
        if isinstance(interaction_matrix, D.DeviceMatrix):
            return interaction_matrix
        if (("csc" in allow and isinstance(interaction_matrix, sp.csc_matrix))
                or ("csr" in allow and isinstance(interaction_matrix, sp.csr_matrix))):
            return D.DeviceMatrix.from_scipy(interaction_matrix)
        raise ValueError(err)

    def _dist_ctx(self):
        """(rank, world) when the SPMD multi-GPU mode is active, else None."""
        if not getattr(self, "distributed", False):
            return None
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
            return None
        return dist.get_rank(), dist.get_world_size()

    def _fit_device(self, X: D.DeviceMatrix, targets: np.ndarray, keep_old: bool, sel_in: Optional[np.ndarray] = None):
        """gram -> solve -> merge.  ``targets``: item ids (host int array)."""
        t = D.require_cuda()
        old = self._device_W() if keep_old else None
        cfg = self._config(X, into_empty_w=(old is None or old.nnz == 0))
        n_items = X.n_items
        tg = D.to_dev(np.ascontiguousarray(targets, dtype=np.int32))
        res = None
        ctx = self._dist_ctx()
        if sel_in is None and not self.keep_fit_details:
            # positive coefficients on non-negative data (all features, or a bulk fit with feature selection): only the Gram
            # rows that can matter are formed when they are few (every rank of an SPMD job does the same small fit: there is
            # nothing worth sharding); None = path does not apply / not worth it
            res = D.fit_pruned(X, tg, cfg)
        if res is None and ctx is not None and X.nnz > 0 and sel_in is None:
            # item-sharded fit: this rank solves the targets of its own Gram row blocks; the solver outputs (a few
            # MB) are all-gathered so that every rank assembles the same full W
            from ... import pipeline as P
            part = P.fit_owner_rows(X, cfg, rank=ctx[0], world=ctx[1], targets=tg if len(targets) != n_items else None)
            if part is not None:
                res = P.gather_solve_results(part, ctx[1])
                targets = res.targets.cpu().numpy()
        if res is None:
            # this G is read by the one solve below: with feature selection in a bulk fit, rows the solver never touches
            # (targets without a live coordinate) are not written back to item order
            live = cfg if (sel_in is None and not self.keep_fit_details) else None
            G = D.gram_full(X, live_cfg=live) if X.nnz > 0 else D.gram(X)
            sel_dev = None
            if sel_in is not None and cfg.nn > 0:
                sel_dev = D.to_dev(np.ascontiguousarray(sel_in, dtype=np.int32).reshape(-1))
            res = D.solve(G, n_items, tg, cfg, sel_in=sel_dev, want_sel=self.keep_fit_details)
            del G
        self._W = D.w_merge(old, n_items, res)
        self._W_host = None
        if self.keep_fit_details:
            self.last_fit_stats = res.stats.cpu().numpy()
            self.last_fit_targets = np.asarray(targets, dtype=np.int64).copy()
            self.last_fit_sel = res.sel.cpu().numpy() if res.sel is not None else None
        return self

    # ------------------------------------------------------------------ fit API
    def fit(self, interaction_matrix, parallel: bool = False, progress_bar: bool = False):
        """All columns (slim_elastic.py:229-281).  ``parallel`` only selects the reference's
        merge semantics: the serial path starts from an empty W, ``fit_in_parallel`` keeps
        entries of an existing W that the new solve does not return (:322-327)."""
        self._check_optim()
        if isinstance(interaction_matrix, sp.csc_matrix) and parallel:
            return self.fit_in_parallel(interaction_matrix, progress_bar=progress_bar)
        if isinstance(interaction_matrix, sp.csr_matrix) and parallel:
            logging.warning("Multiprocessing is only supported for CSC format. Fitting in single process.")
        X = self._as_device(interaction_matrix)
        if isinstance(interaction_matrix, D.DeviceMatrix) and parallel:
            return self._fit_device(X, np.arange(X.n_items), keep_old=True)
        return self._fit_device(X, np.arange(X.n_items), keep_old=False)

    def fit_in_parallel(self, interaction_matrix, item_ids: Optional[ndarray] = None, progress_bar: bool = False,
                        chunk_size: int = 100, num_workers: Optional[int] = None):
        """slim_elastic.py:283-386; chunk_size / num_workers are accepted and ignored (the GPU
        solves all requested columns in one batched launch)."""
        self._check_optim()
        X = self._as_device(interaction_matrix, allow=("csc",),
                            err="Interaction matrix must be in CSC format for parallel processing.")
        if item_ids is None:
            item_ids = np.arange(X.n_items)
        return self._fit_device(X, np.asarray(item_ids), keep_old=True)

    def partial_fit(self, interaction_matrix, user_ids: List[int], parallel: bool = False, progress_bar: bool = False):
        """slim_elastic.py:495-508 (items of the given users are re-solved)."""
        csr = interaction_matrix if isinstance(interaction_matrix, sp.csr_matrix) else None
        if csr is None:
            raise ValueError("Interaction matrix must be a scipy.sparse.csr_matrix.")
        items = set()
        for u in user_ids:
            items.update(csr[u, :].indices.tolist())
        return self.partial_fit_items(interaction_matrix, list(items), progress_bar)

    def partial_fit_items(self, interaction_matrix, updated_items: List[int], parallel: bool = False,
                          progress_bar: bool = False, sel_in: Optional[np.ndarray] = None):
        """slim_elastic.py:510-564: re-solve ``updated_items`` and merge into the existing W."""
        self._check_optim()
        X = self._as_device(interaction_matrix)
        return self._fit_device(X, np.asarray(list(updated_items), dtype=np.int64), keep_old=True, sel_in=sel_in)

    # ------------------------------------------------------------------ scoring API
    def _require_fitted(self, what: str) -> D.DeviceW:
        W = self._device_W()
        if W is None:
            raise RuntimeError(f"Model must be fitted before calling {what}.")
        return W

    def recommend(self, user_id: int, interaction_matrix, candidate_item_ids: Optional[List[int]] = None,
                  top_k: int = 10, filter_interacted: bool = True, dense_output: bool = True, ret_scores: bool = False):
        """slim_elastic.py:628-672."""
        self._require_fitted("predict")
        return self.recommend_batch([user_id], interaction_matrix, candidate_item_ids, top_k, filter_interacted,
                                    dense_output, ret_scores)[0]

    def recommend_batch(self, user_ids: List[int], interaction_matrix, candidate_item_ids: Optional[List[int]] = None,
                        top_k: int = 10, filter_interacted: bool = True, dense_output: bool = True,
                        ret_scores: bool = False):
        """slim_elastic.py:674-741."""
        W = self._require_fitted("batch_recommend")
        if len(user_ids) == 0:
            return []
        X = self._as_device(interaction_matrix, allow=("csr",), err="Interaction matrix must be a scipy.sparse.csr_matrix.")
        ids, scores, cnt = self.recommend_batch_device(np.asarray(user_ids, dtype=np.int64), X, candidate_item_ids,
                                                       top_k, filter_interacted, dense_output)
        out = []
        for r in range(len(user_ids)):
            c = int(cnt[r])
            items = ids[r, :c].tolist()
            if ret_scores:
                if candidate_item_ids is None and not dense_output and c == 0:
                    # zip(*[]) in the reference (slim_elastic.py:815)
                    raise ValueError("not enough values to unpack (expected 2, got 0)")
                out.append((items, scores[r, :c].astype(np.float32)))
            else:
                out.append(items)
        return out

    @staticmethod
    def _check_user_ids(user_ids: np.ndarray, X: D.DeviceMatrix) -> None:
        """The kernels index rptr[user] without a bounds test: reject what the reference would reject (scipy row
        indexing raises IndexError for rows outside the matrix; negative Python indices are not supported here)."""
        if len(user_ids) and (int(user_ids.min()) < 0 or int(user_ids.max()) >= X.n_users):
            raise IndexError(f"user id out of range [0, {X.n_users})")

    @staticmethod
    def _check_item_ids(item_ids: np.ndarray, n_items: int, what: str = "candidate item") -> None:
        if len(item_ids) and (int(item_ids.min()) < 0 or int(item_ids.max()) >= n_items):
            raise IndexError(f"{what} id out of range [0, {n_items})")

    def recommend_batch_device(self, user_ids: np.ndarray, X: D.DeviceMatrix, candidate_item_ids=None, top_k: int = 10,
                               filter_interacted: bool = True, dense_output: bool = True):
        """Array form: returns host arrays (ids int64 [Q,k] with -1 padding, scores f32 [Q,k], cnt [Q])."""
        W = self._require_fitted("batch_recommend")
        n_items = W.n_items
        if X.n_items != n_items:
            # the reference multiplies (n_users x I_x) by (I_w x I_w): shapes must agree
            raise ValueError(f"dimension mismatch: interaction matrix has {X.n_items} items, W has {n_items}")
        user_ids = np.ascontiguousarray(user_ids, dtype=np.int64)
        self._check_user_ids(user_ids, X)
        k = int(top_k)
        if candidate_item_ids is not None:
            cand_host = np.ascontiguousarray(candidate_item_ids, dtype=np.int64)
            self._check_item_ids(cand_host, n_items)
            if min(k, len(cand_host)) > _K_DEVICE:
                return self._topk_large(user_ids, X, k, filter_interacted, dense_output, cand_host)
            users = D.to_dev(user_ids.astype(np.int32))
            cand = D.to_dev(cand_host.astype(np.int32))
            kk = max(1, min(k, len(cand_host)))
            pos, scores, cnt = D.recommend_candidates(X, users, W, cand, kk)
            pos = pos.cpu().numpy().astype(np.int64)
            ids = np.where(pos >= 0, cand_host[np.clip(pos, 0, len(cand_host) - 1)], -1)
            return ids, scores.cpu().numpy(), cnt.cpu().numpy()
        if k > _K_DEVICE:
            return self._topk_large(user_ids, X, k, filter_interacted, dense_output, None)
        users = D.to_dev(user_ids.astype(np.int32))
        mode = RT_TOPK_DENSE if dense_output else RT_TOPK_SPARSE
        ids, scores, cnt = D.recommend(X, users, W, max(1, k), filter_interacted, mode)
        return ids.cpu().numpy().astype(np.int64), scores.cpu().numpy(), cnt.cpu().numpy()

    def _topk_large(self, user_ids: np.ndarray, X: D.DeviceMatrix, k: int, filter_interacted: bool, dense_output: bool,
                    cand_host: Optional[np.ndarray]):
        """top_k > 128 (the fused kernels keep at most 128 entries per user in shared memory): the scores come from the
        candidate-scoring kernel (every requested column with its float32 score, 128 columns per launch), the selection
        follows slim_elastic.py:723-818 on those scores: candidates / dense mode ``argsort(scores)[-k:][::-1]`` minus
        -inf entries, sparse mode = non-zero scores by descending score.  Users are processed in chunks of bounded size."""
        n_items = X.n_items
        k_eff = min(k, len(cand_host) if cand_host is not None else n_items)
        Q = len(user_ids)
        ids = np.full((Q, k_eff), -1, dtype=np.int64)
        scores = np.zeros((Q, k_eff), dtype=np.float32)
        cnt = np.zeros(Q, dtype=np.int32)
        n_cols = len(cand_host) if cand_host is not None else n_items
        step = max(1, min(4096, (1 << 26) // max(n_cols, 1)))
        rptr = ridx = None
        if filter_interacted and cand_host is None:
            rptr, ridx = X.rptr.cpu().numpy(), X.ridx[:X.nnz].cpu().numpy()
        for a in range(0, Q, step):
            S = self._scores_device(user_ids[a:a + step], X, cand_host, "batch_recommend")
            for r in range(S.shape[0]):
                sc = S[r]
                if cand_host is not None:
                    top = np.argsort(sc)[-k:][::-1]
                    out_ids = cand_host[top]
                else:
                    if rptr is not None:
                        u = int(user_ids[a + r])
                        sc[ridx[rptr[u]:rptr[u + 1]]] = -np.inf
                    if dense_output:
                        top = np.argsort(sc)[-k:][::-1]
                        top = top[sc[top] != -np.inf]
                    else:
                        nz = np.flatnonzero((sc != 0) & (sc != -np.inf))
                        top = nz[np.argsort(-sc[nz], kind="stable")][:k]
                    out_ids = top
                c = len(top)
                ids[a + r, :c] = out_ids
                scores[a + r, :c] = sc[top]
                cnt[a + r] = c
        return ids, scores, cnt

    _LIST_CHUNK = 16384  # users per launch when the result is wanted as Python lists

    def recommend_lists(self, user_ids: np.ndarray, X: D.DeviceMatrix, top_k: int = 10, filter_interacted: bool = True,
                        dense_output: bool = True) -> List[List[int]]:
        """``recommend_batch`` without candidates for many users, as Python lists.  The users are scored in
        chunks; every chunk's kernel and its device->host copy (into a pinned buffer) are queued up front, and
        the host turns chunk c into lists while the GPU is still scoring chunk c+1."""
        W = self._require_fitted("batch_recommend")
        t = D.require_cuda()
        if X.n_items != W.n_items:
            raise ValueError(f"dimension mismatch: interaction matrix has {X.n_items} items, W has {W.n_items}")
        Q = len(user_ids)
        user_ids = np.ascontiguousarray(user_ids, dtype=np.int64)
        self._check_user_ids(user_ids, X)
        if int(top_k) > _K_DEVICE:
            ids, _, cnt = self._topk_large(user_ids, X, int(top_k), filter_interacted, dense_output, None)
            return [row[:c].tolist() for row, c in zip(ids, cnt.tolist())]
        k = max(1, int(top_k))
        mode = RT_TOPK_DENSE if dense_output else RT_TOPK_SPARSE
        users = D.to_dev(user_ids.astype(np.int32))
        ctx = self._dist_ctx()
        if ctx is not None and Q >= 64 * ctx[1] and getattr(self, "distributed_queries", "replicated") != "local":
            # query-sharded scoring: every rank scores its slice of the users, the finished lists are all-gathered
            from ... import pipeline as P
            ids, _, cnt = P.recommend_query_sharded(X, users, W, k, filter_interacted, mode, rank=ctx[0], world=ctx[1])
            rows = ids.cpu().numpy().tolist()
            c = cnt.cpu().numpy()
            for r in np.flatnonzero(c < k).tolist():
                del rows[r][int(c[r]):]
            return rows
        pin = _PINNED.get("lists")
        if pin is None or pin[0].numel() < Q * k or pin[1].numel() < Q:
            pin = (t.empty(Q * k, dtype=t.int32, pin_memory=True), t.empty(Q, dtype=t.int32, pin_memory=True))
            _PINNED["lists"] = pin  # grow-only, shared by every model of the process (single caller thread)
        h_ids, h_cnt = pin[0][:Q * k].view(Q, k), pin[1][:Q]
        n_chunks = -(-Q // self._LIST_CHUNK)
        h_redo = t.empty(n_chunks, dtype=t.int32, pin_memory=True)
        pending = []
        for c, a in enumerate(range(0, Q, self._LIST_CHUNK)):
            b = min(a + self._LIST_CHUNK, Q)
            # no host wait inside: the users the tensor-core path hands back are re-scored in a fixed number of slots
            ids, _, cnt, redo = D.recommend(X, users[a:b], W, k, filter_interacted, mode, no_sync=True)
            h_ids[a:b].copy_(ids, non_blocking=True)
            h_cnt[a:b].copy_(cnt, non_blocking=True)
            slots = 0
            if redo is not None:
                h_redo[c:c + 1].copy_(redo[0].view(1), non_blocking=True)
                slots = redo[1]
            ev = t.cuda.Event()
            ev.record()
            pending.append((a, b, ev, c, slots))
        out: List[List[int]] = []
        ids_np, cnt_np = h_ids.numpy(), h_cnt.numpy()
        # a million small lists and ints are created below and none of them can be part of a cycle: keep the
        # cyclic collector from rescanning the growing result while it is being built
        gc_on = gc.isenabled()
        gc.disable()
        try:
            for a, b, ev, ci, slots in pending:
                ev.synchronize()
                if slots and int(h_redo[ci]) > slots:
                    # more hand-backs than slots (not seen on the benchmark shapes): this chunk again, waiting for the list
                    ids, _, cnt = D.recommend(X, users[a:b], W, k, filter_interacted, mode)
                    ids_np[a:b] = ids.cpu().numpy()
                    cnt_np[a:b] = cnt.cpu().numpy()
                rows = ids_np[a:b].tolist()
                c = cnt_np[a:b]
                for r in np.flatnonzero(c < k).tolist():  # lists shorter than k: drop the -1 padding
                    del rows[r][int(c[r]):]
                out.extend(rows)
        finally:
            if gc_on:
                gc.enable()
        return out

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_W_host"] = self.item_similarity
        st["_W"] = None
        return st

    def similar_items(self, item_id: int, top_k: int = 10, ret_ndarrays: bool = False):
        """slim_elastic.py:820-857."""
        W = self._require_fitted("similar_items")
        ids, scores, cnt = self.similar_items_batch(np.asarray([item_id]), top_k)
        c = int(cnt[0])
        if ret_ndarrays:
            return ids[0, :c].astype(np.int32), scores[0, :c]
        return list(zip(ids[0, :c].tolist(), scores[0, :c].tolist()))

    def similar_items_batch(self, item_ids: np.ndarray, top_k: int = 10):
        W = self._require_fitted("similar_items")
        item_ids = np.ascontiguousarray(item_ids, dtype=np.int64)
        self._check_item_ids(item_ids, W.n_items, "query item")   # the reference's W[:, item_id] raises IndexError
        items = D.to_dev(item_ids.astype(np.int32))
        ids, scores, cnt = D.similar(W, items, max(1, int(top_k)))
        return ids.cpu().numpy().astype(np.int64), scores.cpu().numpy(), cnt.cpu().numpy()

    # ------------------------------------------------------------------ dense score export (slim_elastic.py:566-626)
    # Not called by SLIM / Recommender (nor by anything else in the reference package); served by the candidate
    # scoring kernel (rt_slim_recommend_candidates: every candidate is returned with its score), 128 columns per launch.
    def _scores_device(self, user_ids: np.ndarray, X: D.DeviceMatrix, item_ids=None, what: str = "predict") -> np.ndarray:
        W = self._require_fitted(what)
        if X.n_items != W.n_items:
            raise ValueError(f"dimension mismatch: interaction matrix has {X.n_items} items, W has {W.n_items}")
        user_ids = np.ascontiguousarray(user_ids, dtype=np.int64)
        if len(user_ids) and (user_ids.min() < 0 or user_ids.max() >= X.n_users):
            raise IndexError("row index out of range")
        cols = np.arange(W.n_items, dtype=np.int32) if item_ids is None else np.ascontiguousarray(item_ids, dtype=np.int32)
        if len(cols) and (cols.min() < 0 or cols.max() >= W.n_items):
            raise IndexError("column index out of range")
        Q = len(user_ids)
        out = np.zeros((Q, len(cols)), dtype=np.float32)
        if Q == 0 or len(cols) == 0:
            return out
        users = D.to_dev(user_ids.astype(np.int32))
        rows = np.arange(Q)[:, None]
        for a in range(0, len(cols), 128):
            chunk = cols[a:a + 128]
            pos, sc, cnt = D.recommend_candidates(X, users, W, D.to_dev(chunk), len(chunk))
            pos, sc, cnt = pos.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()
            assert (cnt == len(chunk)).all(), "candidate scoring returns every candidate"
            out[rows, a + pos] = sc
        return out

    @staticmethod
    def _shape_scores(S: np.ndarray, dense_output: bool):
        return S if dense_output else sp.csr_matrix(S)   # exact zeros are not stored, like scipy's csr_matmat result

    def predict(self, user_id: int, interaction_matrix, dense_output: bool = True):
        """slim_elastic.py:566-586: scores of one user over all items, shape (1, n_items)."""
        self._require_fitted("predict")
        X = self._as_device(interaction_matrix, allow=("csr",), err="Interaction matrix must be a scipy.sparse.csr_matrix.")
        return self._shape_scores(self._scores_device(np.asarray([user_id]), X, None, "predict"), dense_output)

    def predict_selected(self, user_id: int, item_ids: List[int], interaction_matrix, dense_output: bool = True):
        """slim_elastic.py:588-608: scores of one user over ``item_ids``, shape (1, len(item_ids))."""
        self._require_fitted("predict_selected")
        X = self._as_device(interaction_matrix, allow=("csr",), err="Interaction matrix must be a scipy.sparse.csr_matrix.")
        return self._shape_scores(self._scores_device(np.asarray([user_id]), X, item_ids, "predict_selected"), dense_output)

    def predict_all(self, interaction_matrix, dense_output: bool = True):
        """slim_elastic.py:610-626: scores of every user over all items, shape (n_users, n_items) -- a dense matrix, as in
        the reference; meant for small matrices."""
        self._require_fitted("predict_all")
        X = self._as_device(interaction_matrix, allow=("csr",), err="Interaction matrix must be a scipy.sparse.csr_matrix.")
        return self._shape_scores(self._scores_device(np.arange(X.n_users), X, None, "predict_all"), dense_output)
