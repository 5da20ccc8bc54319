"""Minimal tag store kept only because ``BaseModel`` instantiates one and the reference's SLIM tests
register tags through it (/root/reference/rtrec/models/base.py:34-70,
/root/reference/tests/models/test_slim.py:11-27).  The SLIM hot path never reads it; the
LightFM/Hybrid consumers of the reference's FeatureStore are out of scope.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
from scipy.sparse import csr_matrix


class _Side:
    def __init__(self) -> None:
        self.vocab: Dict[str, int] = {}
        self.rows: Dict[int, List[int]] = {}

    def put(self, row_id: int, tags: List[str], append: bool) -> None:
        cur = list(self.rows.get(row_id, [])) if append else []
        for tag in tags:
            tid = self.vocab.setdefault(tag, len(self.vocab))
            if tid not in cur:
                cur.append(tid)
        self.rows[row_id] = cur

    def clear(self, ids: Optional[List[int]]) -> None:
        if ids is None:
            self.rows.clear()
        else:
            for i in ids:
                self.rows.pop(i, None)

    def matrix(self, ids: Optional[List[int]]) -> csr_matrix:
        if ids is None:
            n_rows = (max(self.rows) + 1) if self.rows else 0
            ids = list(range(n_rows))
        indptr, cols = [0], []
        for i in ids:
            cols.extend(self.rows.get(i, []))
            indptr.append(len(cols))
        data = np.ones(len(cols), dtype=np.float32)
        return csr_matrix((data, np.asarray(cols, dtype=np.int32), np.asarray(indptr, dtype=np.int32)),
                          shape=(len(ids), len(self.vocab)))


class FeatureStore:
    def __init__(self) -> None:
        self._users = _Side()
        self._items = _Side()

    def num_user_features(self) -> int:
        return len(self._users.vocab)

    def num_item_features(self) -> int:
        return len(self._items.vocab)

    def put_user_features(self, user_id: int, user_tags: List[str], append: bool = False) -> None:
        self._users.put(user_id, user_tags, append)

    def put_item_features(self, item_id: int, item_tags: List[str], append: bool = False) -> None:
        self._items.put(item_id, item_tags, append)

    def clear_user_features(self, user_ids: Optional[List[int]] = None) -> None:
        self._users.clear(user_ids)

    def clear_item_features(self, item_ids: Optional[List[int]] = None) -> None:
        self._items.clear(item_ids)

    def build_user_features_matrix(self, user_ids: Optional[List[int]] = None) -> csr_matrix:
        return self._users.matrix(user_ids)

    def build_item_features_matrix(self, item_ids: Optional[List[int]] = None) -> csr_matrix:
        return self._items.matrix(item_ids)
