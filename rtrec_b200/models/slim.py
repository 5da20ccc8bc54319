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
