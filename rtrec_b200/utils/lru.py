"""Recency-ordered frequency set used for cold-start "hot items".

Behavioural mirror of the reference's ``LRUFreqSet`` (/root/reference/rtrec/utils/lru.py:11-123):
keys are kept in least-recently-used -> most-recently-used order with a hit counter; when the set
is full the least recently used key is evicted; ``get_freq_items`` lists keys by counter
(descending, ties in recency order).  ``add_batch`` is the vectorised equivalent of calling
``add`` for every element of an array in order (used by the batched ingest path); it must and does
produce exactly the same state.
"""
from __future__ import annotations

from collections import OrderedDict
from collections.abc import MutableSet
from typing import Any, Iterator, List, Optional

import numpy as np


class LRUFreqSet(MutableSet):
    def __init__(self, capacity: int):
        if capacity <= 0:
            raise ValueError("Capacity must be greater than 0.")
        self.capacity = capacity
        self.data: "OrderedDict[Any, int]" = OrderedDict()

    # -- MutableSet protocol -----------------------------------------------------------------
    def add(self, value: Any) -> None:
        hits = self.data.pop(value, None)
        if hits is None:
            if len(self.data) >= self.capacity:
                self.data.popitem(last=False)
            hits = 0
        self.data[value] = hits + 1

    def discard(self, value: Any) -> None:
        if value not in self.data:
            raise KeyError(value)
        del self.data[value]

    def __contains__(self, key: Any) -> bool:
        return key in self.data

    def __iter__(self) -> Iterator[Any]:
        return iter(self.data)

    def __len__(self) -> int:
        return len(self.data)

    def __repr__(self) -> str:
        return f"LRUFreqSet(capacity={self.capacity}, size={len(self.data)})"

    # -- batched update ----------------------------------------------------------------------
    def add_batch(self, values: np.ndarray) -> None:
        """Same end state as ``for v in values: self.add(v)``."""
        n = len(values)
        if n == 0:
            return
        values = np.asarray(values)
        if values.dtype.kind in "iu" and values.min() >= 0 and int(values.max()) < 8 * n + (1 << 20):
            # small non-negative ints: O(n) counting instead of a sort
            dense_counts = np.bincount(values)
            uniq = np.flatnonzero(dense_counts)
            counts = dense_counts[uniq]
            dense_last = np.zeros(len(dense_counts), dtype=np.int64)
            dense_last[values] = np.arange(n)  # repeated index: the last assignment wins
            last = dense_last[uniq]
        else:
            uniq, inv, counts = np.unique(values, return_inverse=True, return_counts=True)
            last = np.full(len(uniq), -1, dtype=np.int64)
            last[inv] = np.arange(n)  # later occurrences overwrite earlier ones
        if not self.add_counts(uniq, counts, last):
            # an eviction can happen somewhere inside the batch: replay it event by event
            if not self._replay_native(values):
                for v in values.tolist():
                    self.add(v)

    def _replay_native(self, values: np.ndarray) -> bool:
        """Event-by-event replay in the library's host code (``rt_lru_replay``) for non-negative integer keys; False when
        the keys do not qualify (the caller then runs the Python loop)."""
        if values.dtype.kind not in "iu" or len(values) < 4096:
            return False
        try:
            import ctypes as C
            from .. import _lib
            lib = _lib.load()
            cur_k = np.fromiter(self.data.keys(), dtype=np.int64, count=len(self.data))
            cur_c = np.fromiter(self.data.values(), dtype=np.int64, count=len(self.data))
        except Exception:  # noqa: BLE001 - library not built, or non-integer keys in the set
            return False
        lo = min(int(values.min()), int(cur_k.min()) if len(cur_k) else 0)
        bound = max(int(values.max()), int(cur_k.max()) if len(cur_k) else 0) + 1
        if lo < 0 or bound > (1 << 28):
            return False
        vals = np.ascontiguousarray(values, dtype=np.int64)
        out_k = np.empty(self.capacity, dtype=np.int64)
        out_c = np.empty(self.capacity, dtype=np.int64)
        n_out = C.c_int64(0)
        rc = lib.rt_lru_replay(vals.ctypes.data, len(vals), self.capacity, bound, cur_k.ctypes.data, cur_c.ctypes.data, len(cur_k),
                               out_k.ctypes.data, out_c.ctypes.data, C.byref(n_out))
        if rc != 0:
            return False
        m = int(n_out.value)
        self.data = OrderedDict(zip(out_k[:m].tolist(), out_c[:m].tolist()))
        return True

    def add_counts(self, uniq: np.ndarray, counts: np.ndarray, last: np.ndarray) -> bool:
        """Apply a batch summarised as (distinct key, number of adds, arrival index of its last add).
        Same end state as the event-by-event loop provided no eviction can occur; returns False
        (and changes nothing) when one could, so the caller replays the events one by one."""
        data = self.data
        if not data:
            # empty set (first batch of a bulk fit): the result is simply the keys in last-touch order
            if len(uniq) > self.capacity:
                return False
            order = np.argsort(last, kind="stable")
            data.update(zip(np.asarray(uniq)[order].tolist(), np.asarray(counts)[order].tolist()))
            return True
        keys = uniq.tolist()
        n_fresh = len(keys) - sum(map(data.__contains__, keys))
        if len(data) + n_fresh > self.capacity:
            return False
        # no eviction: counters add up, touched keys move to the MRU end ordered by last touch
        cnt = counts.tolist()
        for j in np.argsort(last, kind="stable").tolist():
            key = keys[j]
            data[key] = data.pop(key, 0) + cnt[j]
        return True

    # -- queries -----------------------------------------------------------------------------
    def get_freq_items(self, n: Optional[int] = None, exclude_items: List[Any] = []) -> Iterator[Any]:
        ranked = sorted(self.data.items(), key=lambda kv: kv[1], reverse=True)  # stable: recency breaks ties
        if len(exclude_items) > 0:
            emitted = 0
            for key, _ in ranked:
                if key in exclude_items:
                    continue
                if n is not None and emitted >= n:
                    break
                yield key
                emitted += 1
        else:
            for key, _ in (ranked if n is None else ranked[:n]):
                yield key
