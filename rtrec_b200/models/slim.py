"""``SLIM`` -- glue between the device store and the device operator, behind the reference's model API.

Mirror of /root/reference/rtrec/models/slim.py:21-149.  Differences are only in where the data
lives: matrices handed to the operator are :class:`DeviceMatrix` objects built on the GPU from
the device store (no scipy round trip), and ``recommend_batch`` scores a whole list of users in
one fused kernel launch.
"""
from __future__ import annotations

import logging
from typing import Any, Iterable, List, Optional, Tuple

import numpy as np

from .base import BaseModel
from .internal.slim_elastic import SLIMElastic


class SLIM(BaseModel):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.model = SLIMElastic(kwargs)
        self.recorded_item_ids = set()

    # ------------------------------------------------------------------ fitting
    def fit(self, interactions: Iterable[Tuple[Any, Any, float, float]], update_interaction: bool = False,
            progress_bar: bool = True):
        """slim.py:28-43: ingest, then re-solve exactly the item columns seen in this call on a matrix
        that holds only those columns."""
        seen = _RecordingSet()
        prev, self.recorded_item_ids = self.recorded_item_ids, seen
        try:
            self.add_interactions(interactions, update_interaction=update_interaction, record_interactions=True)
        finally:
            self.recorded_item_ids = prev
        item_ids = list(seen)
        X = self.interactions.device_matrix(item_ids)
        self.model.partial_fit_items(X, item_ids, progress_bar=progress_bar)
        return self

    def _record_interactions(self, user_id: int, item_id: int, tstamp: float, rating: float) -> None:
        self.recorded_item_ids.add(item_id)

    def _record_interaction_arrays(self, user_ids: np.ndarray, item_ids: np.ndarray) -> None:
        self.recorded_item_ids.update(np.unique(item_ids).tolist())

    def _fit_recorded(self, parallel: bool = False, progress_bar: bool = True):
        """slim.py:48-53."""
        item_ids = list(self.recorded_item_ids)
        X = self.interactions.device_matrix(item_ids)
        self.model.partial_fit_items(X, item_ids, parallel=parallel, progress_bar=progress_bar)
        self.recorded_item_ids.clear()
        return self

    def bulk_fit(self, parallel: bool = False, progress_bar: bool = True):
        """slim.py:55-64."""
        X = self.interactions.device_matrix()
        self.model.fit(X, parallel=parallel, progress_bar=progress_bar)
        return self

    # ------------------------------------------------------------------ scoring
    def _recommend(self, user_id: int, candidate_item_ids: Optional[List[int]] = None, user_tags: Optional[List[str]] = None,
                   top_k: int = 10, filter_interacted: bool = True) -> List[int]:
        """slim.py:66-79."""
        return self._recommend_hot_batch([user_id], candidate_item_ids=candidate_item_ids, top_k=top_k,
                                         filter_interacted=filter_interacted)[0]

    def _recommend_hot_batch(self, user_ids: List[int], candidate_item_ids: Optional[List[int]] = None,
                             users_tags: Optional[List[List[str]]] = None, top_k: int = 10,
                             filter_interacted: bool = True) -> List[List[int]]:
        """slim.py:81-104: ``dense_output`` follows the id kind (ints -> sparse top-k semantics)."""
        if self.model._W is None and self.model._W_host is None:  # (not .item_similarity: that would download W)
            raise RuntimeError("Model must be fitted before calling batch_recommend.")
        if len(user_ids) == 0:
            return []
        X = self.interactions.device_matrix()
        dense_output = not self.item_ids.pass_through
        if candidate_item_ids is None:
            return self.model.recommend_lists(np.asarray(user_ids, dtype=np.int64), X, top_k, filter_interacted, dense_output)
        ids, _, cnt = self.model.recommend_batch_device(np.asarray(user_ids, dtype=np.int64), X, candidate_item_ids,
                                                        top_k, filter_interacted, dense_output)
        rows = ids.tolist()
        if len(cnt) and int(cnt.min()) == ids.shape[1]:
            return rows  # every list is full: nothing to trim
        return [row if c == len(row) else row[:c] for row, c in zip(rows, cnt.tolist())]

    def _similar_items(self, query_item_id: int, query_item_tags: Optional[List[str]] = None, top_k: int = 10) -> List[Tuple[int, float]]:
        """slim.py:106-115."""
        return self.model.similar_items(query_item_id, top_k=top_k, ret_ndarrays=False)  # type: ignore

    def similar_items_batch(self, query_items: List[Any], top_k: int = 10, ret_scores: bool = False) -> List[list]:
        """``[self.similar_items(q, None, top_k, ret_scores) for q in query_items]`` (recommender.py:150-161 ->
        base.py:320-340) with ONE kernel launch and one device->host copy for the whole query list."""
        if len(query_items) == 0:
            return []
        ids = self.item_ids.identify_many(query_items)   # registers unknown query items like ``identify`` does
        out_ids, scores, cnt = self.model.similar_items_batch(ids, top_k)
        rows, srows, cnts = out_ids.tolist(), scores.tolist(), cnt.tolist()
        to_item = (lambda x: x) if self.item_ids.pass_through else self.item_ids.get
        if ret_scores:
            return [[(to_item(i), sc) for i, sc in zip(r[:c], sr[:c])] for r, sr, c in zip(rows, srows, cnts)]
        return [[to_item(i) for i in r[:c]] for r, c in zip(rows, cnts)]

    def evaluate_device(self, users, gptr: np.ndarray, gt_items, recommend_size: int = 10,
                        filter_interacted: bool = True) -> Optional[dict]:
        """``compute_scores`` of ``Recommender.evaluate`` (recommender.py:163-200, metrics.py:267-313) without Python
        lists: the top-k lists stay on the device and ``rt_eval_metrics`` produces the per-user values, each bit-identical
        to the Python function; they are added up in user order like the reference's ``+=``.  ``users[q]``'s ground truth
        is ``gt_items[gptr[q]:gptr[q+1]]`` (original item objects, duplicates kept).  Returns ``None`` when this path
        does not apply (k > 128): the caller then runs the list-based loop."""
        import sys
        from .. import _lib
        from .. import device as D
        from .._lib import RT_TOPK_DENSE, RT_TOPK_SPARSE
        k = int(recommend_size)
        if k < 1 or k > 128:
            return None
        if self.model._W is None and self.model._W_host is None:
            raise RuntimeError("Model must be fitted before calling batch_recommend.")
        Q = len(users)
        if Q == 0:
            return None
        t = D.require_cuda()
        # ---- users: known -> row of X, unknown -> cold start (hot items, base.py:249-269)
        arr = np.asarray(users)
        if self.user_ids.pass_through and arr.dtype.kind in "iu":
            uid = arr.astype(np.int64)
            uid[(uid < 0) | (uid > self.interactions.max_user_id)] = -1
        else:
            uid = np.asarray([-1 if (r := self._resolve_user(u)) is None else r for u in users], dtype=np.int64)
        # ---- ground truth: item objects -> ids (unknown items never match a recommendation but count in the length)
        gi = np.asarray(gt_items)
        if self.item_ids.pass_through and gi.dtype.kind in "iu":
            gid = gi.astype(np.int64)
            gid[(gid < 0) | (gid > 2**31 - 2)] = -1
        else:
            get_id = self.item_ids.obj_to_id.get if not self.item_ids.pass_through else None
            if get_id is None:
                return None   # integer pass-through ids judged against non-integer ground truth: list path decides
            gid = np.asarray([get_id(x, -1) for x in (gt_items.tolist() if hasattr(gt_items, "tolist") else gt_items)], dtype=np.int64)
        gptr = np.ascontiguousarray(gptr, dtype=np.int64)
        row = np.repeat(np.arange(Q, dtype=np.int64), np.diff(gptr))
        order = np.lexsort((gid, row))
        d_gidx = D.to_dev(gid[order].astype(np.int32)) if len(gid) else D.zeros(1, t.int32)
        d_gptr = D.to_dev(gptr)
        X = self.interactions.device_matrix()
        W = self.model._require_fitted("batch_recommend")
        mode = RT_TOPK_SPARSE if self.item_ids.pass_through else RT_TOPK_DENSE
        hot = np.flatnonzero(uid >= 0)
        ids = D.empty(Q * k, t.int32).view(Q, k)
        cnt = D.zeros(Q, t.int32)
        if len(hot):
            hid, _, hcnt = D.recommend(X, D.to_dev(uid[hot].astype(np.int32)), W, k, filter_interacted, mode)
            if len(hot) == Q:
                ids, cnt = hid, hcnt
            else:
                hpos = D.to_dev(hot)
                ids[hpos] = hid
                cnt[hpos] = hcnt
        if len(hot) < Q:
            cold_list = self.interactions.get_hot_items(k, filter_interacted=False)[:k]
            cpos = D.to_dev(np.flatnonzero(uid < 0))
            if cold_list:
                ids[cpos, :len(cold_list)] = D.to_dev(np.asarray(cold_list, dtype=np.int32))
            cnt[cpos] = len(cold_list)
        from math import log2
        disc = D.to_dev(np.asarray([1 / log2(i + 2) for i in range(max(k, 1))], dtype=np.float64))
        out = D.empty(Q * 9, t.float64)
        _lib.check(_lib.load().rt_eval_metrics(D.ptr(ids.contiguous()), D.ptr(cnt.contiguous()), Q, k, k, D.ptr(d_gptr),
                                               D.ptr(d_gidx), D.ptr(disc), 1 if sys.version_info >= (3, 12) else 0,
                                               D.ptr(out), D.stream_ptr()), "rt_eval_metrics")
        per_user = out.view(Q, 9).cpu().numpy()
        sums = np.cumsum(per_user, axis=0)[-1]   # strictly sequential float64 additions = the reference's `+=` per user
        names = ("precision", "recall", "f1", "ndcg", "hit_rate", "mrr", "map", "tp", "auc")
        res = {n: float(sums[c]) / Q for c, n in enumerate(names)}
        res["tp"] = int(per_user[:, 7].sum())
        return res

    # ------------------------------------------------------------------ persistence (slim.py:117-149)
    def _serialize(self) -> dict:
        return {
            "model": self.model,
            "interactions": self.interactions,
            "user_ids": self.user_ids,
            "item_ids": self.item_ids,
            "feature_store": self.feature_store,
        }

    @classmethod
    def _deserialize(cls, data: dict):
        instance = cls()
        instance.model = data["model"]
        instance.interactions = data["interactions"]
        instance.user_ids = data["user_ids"]
        instance.item_ids = data["item_ids"]
        instance.feature_store = data["feature_store"]
        return instance


class _RecordingSet(set):
    """item ids recorded during one ``fit`` call (kept apart from ``recorded_item_ids`` like the
    reference's local ``item_id_set``, slim.py:29)."""
