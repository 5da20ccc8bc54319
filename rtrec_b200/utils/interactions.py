"""Device-resident user-item interaction store behind the reference's ``UserItemInteractions`` API.

Reference: /root/reference/rtrec/utils/interactions.py:14-353 (dict-of-dicts, python loops).  Here
the truth lives in HBM as three arrays sorted by ``(user << 32 | item)`` (stored rating f64, last
timestamp f64); events are buffered on the host in arrival order and folded in by the K1 kernels
(``rt_store_fold``) the next time anything reads the store, and the float32 CSR/CSC matrices are
produced by the fused decay+cast+build kernels (``rt_store_build``).  Host-side bookkeeping that
the reference keeps exact (``max_user_id``, ``max_item_id``, ``max_timestamp``, ``all_item_ids``,
``hot_items``) stays on the host and is updated per event / per batch with identical results.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import time
from datetime import datetime, timezone
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .. import _lib
from .. import device as D
from .lru import LRUFreqSet

_INT32_MAX = 2**31 - 1
_FLUSH_AT = 1 << 22  # pending events before an automatic fold
_DEVICE_INGEST_AT = 1 << 16  # batch size from which the bookkeeping runs on the device


class UserItemInteractions:
    def __init__(self, min_value: int = -5, max_value: int = 10, decay_in_days: Optional[int] = None, **kwargs: Any) -> None:
        n_recent_hot = kwargs.get("n_recent_hot", 100_000)
        self.hot_items = LRUFreqSet(capacity=n_recent_hot)
        assert max_value > min_value, f"max_value should be greater than min_value {max_value} > {min_value}"
        self.min_value = min_value
        self.max_value = max_value
        # half-life decay, interactions.py:36-39
        self.decay_rate = None if decay_in_days is None else 1.0 - (math.log(2) / decay_in_days)
        self.all_item_ids: set[int] = set()
        self.max_user_id = 0
        self.max_item_id = 0
        self.max_timestamp = 0.0
        # SPMD multi-GPU mode (see SLIMElastic.distributed): every rank is handed the same event batch, so each rank
        # uploads only its 1/N slice over its own PCIe link and the slices are all-gathered over NVLink
        self.distributed = bool(kwargs.get("distributed", False))
        # device state (torch tensors) -- sorted by key
        self._keys = None
        self._vals = None
        self._stamps = None
        self._n_pairs = 0
        self._dev_max_ts = 0.0
        # host state used only to carry a pickled store until a GPU is needed
        self._host_state: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None
        # pending events, arrival order
        self._pend: List[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = []
        self._pend_scalar: List[Tuple[int, int, float, float]] = []
        self._pend_upsert: Optional[bool] = None
        self._pend_n = 0
        self._matrix_cache: Dict[Any, D.DeviceMatrix] = {}
        self.version = 0  # bumped on every ingest; lets callers cache derived device data

    # ------------------------------------------------------------------ decay helpers
    def get_decay_rate(self) -> Optional[float]:
        return self.decay_rate

    def set_decay_rate(self, decay_rate: Optional[float]) -> None:
        self.decay_rate = decay_rate
        self._matrix_cache.clear()

    def _apply_decay(self, value: float, last_timestamp: float) -> float:
        if self.decay_rate is None:
            return value
        elapsed_days = (self.max_timestamp - last_timestamp) / 86400.0
        return value * self.decay_rate ** elapsed_days

    # ------------------------------------------------------------------ ingest
    def _warn_future(self, tstamp: float) -> None:
        now = time.time()
        if tstamp > now + 180.0:
            cur = datetime.fromtimestamp(now, tz=timezone.utc).isoformat() + "Z"
            ts = datetime.fromtimestamp(tstamp, tz=timezone.utc).isoformat() + "Z"
            logging.warning(f"Timestamp {ts} is in the future. Current time is {cur}")

    @staticmethod
    def _check_id(v: int, what: str) -> int:
        v = int(v)
        if v < 0 or v > _INT32_MAX:
            raise ValueError(f"{what} id {v} is outside [0, 2^31): the device store indexes with int32")
        return v

    def add_interaction(self, user_id: int, item_id: int, tstamp: float, delta: float = 1.0, upsert: bool = False) -> None:
        """One event (interactions.py:81-119)."""
        user_id = self._check_id(user_id, "user")
        item_id = self._check_id(item_id, "item")
        tstamp = float(tstamp)
        delta = float(delta)
        self._warn_future(tstamp)
        self.max_timestamp = max(self.max_timestamp, tstamp + 1.0)
        if self._pend_upsert is not None and self._pend_upsert != bool(upsert):
            self._flush()
        self._pend_upsert = bool(upsert)
        self._pend_scalar.append((user_id, item_id, tstamp, delta))
        self._pend_n += 1
        self.all_item_ids.add(item_id)
        if delta > 0:
            self.hot_items.add(item_id)
        self.max_user_id = max(self.max_user_id, user_id)
        self.max_item_id = max(self.max_item_id, item_id)
        self._touch()
        if self._pend_n >= _FLUSH_AT:
            self._flush()

    def add_interactions_batch(self, user_ids, item_ids, tstamps, deltas, upsert: bool = False) -> None:
        """Vectorised ``add_interaction`` over arrays in arrival order; same end state."""
        u = np.ascontiguousarray(user_ids, dtype=np.int64)
        i = np.ascontiguousarray(item_ids, dtype=np.int64)
        ts = np.ascontiguousarray(tstamps, dtype=np.float64)
        d = np.ascontiguousarray(deltas, dtype=np.float64)
        n = len(u)
        if n == 0:
            return
        if not (len(i) == n and len(ts) == n and len(d) == n):
            raise ValueError("event arrays must have equal length")
        if n >= _DEVICE_INGEST_AT and self._ingest_on_device(u, i, ts, d, bool(upsert)):
            return
        if u.min() < 0 or i.min() < 0 or u.max() > _INT32_MAX or i.max() > _INT32_MAX:
            raise ValueError("ids outside [0, 2^31): the device store indexes with int32")
        self._warn_future(float(ts.max()))
        self.max_timestamp = max(self.max_timestamp, float(ts.max()) + 1.0)
        if self._pend_upsert is not None and self._pend_upsert != bool(upsert):
            self._flush()
        self._pend_upsert = bool(upsert)
        self._seal_scalars()
        self._pend.append((u.astype(np.int32), i.astype(np.int32), ts, d))
        self._pend_n += n
        imax = int(i.max())
        if imax < 8 * n + (1 << 20):
            self.all_item_ids.update(np.flatnonzero(np.bincount(i)).tolist())
        else:
            self.all_item_ids.update(np.unique(i).tolist())
        pos = d > 0
        self.hot_items.add_batch(i[pos] if not pos.all() else i)
        self.max_user_id = max(self.max_user_id, int(u.max()))
        self.max_item_id = max(self.max_item_id, imax)
        self._touch()
        if self._pend_n >= _FLUSH_AT:
            self._flush()

    def _ingest_on_device(self, u: np.ndarray, i: np.ndarray, ts: np.ndarray, d: np.ndarray, upsert: bool) -> bool:
        """Large batches: upload the four columns as they are and let the device produce the batch
        bookkeeping (id ranges, max timestamp, per-item hot counts / last touch / seen flags); the
        host only walks the distinct items.  When the item ids are too sparse for dense per-item counters, or an
        LRU eviction could occur inside the batch, that part of the bookkeeping is replayed on the host."""
        t = D.require_cuda()
        lib = _lib.load()
        n = len(u)
        # pageable host columns -> device through the library's threaded staging pipeline (ids narrowed to int32)
        lo_u, hi_u, lo_i, hi_i, mx_ts = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_double(0)
        ctx = self._dist_ctx()
        if ctx is None:
            du, di = D.empty(n, t.int32), D.empty(n, t.int32)
            dts, dd = D.empty(n, t.float64), D.empty(n, t.float64)
            _lib.check(lib.rt_upload_events(u.ctypes.data, i.ctypes.data, ts.ctypes.data, d.ctypes.data, n, D.ptr(du), D.ptr(di),
                                            D.ptr(dts), D.ptr(dd), C.byref(lo_u), C.byref(hi_u), C.byref(lo_i), C.byref(hi_i),
                                            C.byref(mx_ts), 0, D.stream_ptr()), "rt_upload_events")
        else:
            # this rank's slice [a, b) of the batch goes up its own PCIe link; NCCL all-gathers the slices and the id
            # ranges / max timestamp (the host never touches the other N-1 slices)
            import torch.distributed as dist
            rank, world = ctx
            per = -(-n // world)
            a, b = min(n, rank * per), min(n, (rank + 1) * per)
            m = b - a
            su, si = D.zeros(per, t.int32), D.zeros(per, t.int32)
            sts, sd = D.zeros(per, t.float64), D.zeros(per, t.float64)
            if m > 0:
                _lib.check(lib.rt_upload_events(u[a:b].ctypes.data, i[a:b].ctypes.data, ts[a:b].ctypes.data, d[a:b].ctypes.data,
                                                m, D.ptr(su), D.ptr(si), D.ptr(sts), D.ptr(sd), C.byref(lo_u), C.byref(hi_u),
                                                C.byref(lo_i), C.byref(hi_i), C.byref(mx_ts), 0, D.stream_ptr()),
                           "rt_upload_events")
            big = float(1 << 40)
            ext = t.tensor([-(lo_u.value if m else big), hi_u.value if m else -big, -(lo_i.value if m else big),
                            hi_i.value if m else -big, mx_ts.value if m else -big], dtype=t.float64, device=D.dev())
            dist.all_reduce(ext, op=dist.ReduceOp.MAX)
            fu, fi = D.empty(per * world, t.int32), D.empty(per * world, t.int32)
            fts, fd = D.empty(per * world, t.float64), D.empty(per * world, t.float64)
            for full, part in ((fu, su), (fi, si), (fts, sts), (fd, sd)):
                dist.all_gather_into_tensor(full, part)
            du, di, dts, dd = fu[:n], fi[:n], fts[:n], fd[:n]
            e = ext.tolist()
            lo_u.value, hi_u.value, lo_i.value, hi_i.value, mx_ts.value = int(-e[0]), int(e[1]), int(-e[2]), int(e[3]), e[4]
        if lo_u.value < 0 or lo_i.value < 0 or hi_u.value > _INT32_MAX or hi_i.value > _INT32_MAX:
            raise ValueError("ids outside [0, 2^31): the device store indexes with int32")
        imax = int(hi_i.value)
        if imax >= 8 * n + (1 << 20):
            return self._queue_device_batch(du, di, dts, dd, upsert, int(hi_u.value), imax, float(mx_ts.value), i, d)
        cnt = D.empty(imax + 1, t.int32); last = D.empty(imax + 1, t.int32); seen = D.empty(imax + 1, t.uint8)
        _lib.check(lib.rt_events_item_stats32(D.ptr(di), D.ptr(dd), n, imax + 1, D.ptr(cnt), D.ptr(last), D.ptr(seen),
                                              D.stream_ptr()), "rt_events_item_stats32")
        cnt_h, last_h, seen_h = cnt.cpu().numpy(), last.cpu().numpy(), seen.cpu().numpy()
        hot = np.flatnonzero(cnt_h)
        if not self.hot_items.add_counts(hot, cnt_h[hot], last_h[hot]):
            # an LRU eviction can happen inside the batch: replay the hot-item updates event by event on the host
            pos = d > 0
            vals = i[pos] if not pos.all() else i
            if not self.hot_items._replay_native(vals):
                self.hot_items.add_batch(vals)
        self.all_item_ids.update(np.flatnonzero(seen_h).tolist())
        return self._queue_device_batch(du, di, dts, dd, upsert, int(hi_u.value), imax, float(mx_ts.value), None, None)

    def _dist_ctx(self):
        """(rank, world) when the SPMD multi-GPU mode is active, else None."""
        if not getattr(self, "distributed", False):
            return None
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
            return None
        return dist.get_rank(), dist.get_world_size()

    def _queue_device_batch(self, du, di, dts, dd, upsert: bool, umax: int, imax: int, ts_max: float,
                            host_items: Optional[np.ndarray], host_delta: Optional[np.ndarray]) -> bool:
        """Common tail of the device ingest: queue the resident columns for the next fold and update the
        exact host bookkeeping.  ``host_items`` given = item ids too sparse for dense device counters:
        all_item_ids / hot_items are then updated from the host arrays."""
        if host_items is not None:
            self.all_item_ids.update(np.unique(host_items).tolist())
            pos = host_delta > 0
            self.hot_items.add_batch(host_items[pos] if not pos.all() else host_items)
        self._warn_future(ts_max)
        self.max_timestamp = max(self.max_timestamp, ts_max + 1.0)
        if self._pend_upsert is not None and self._pend_upsert != upsert:
            self._flush()
        self._pend_upsert = upsert
        self._seal_scalars()
        self._pend.append((du, di, dts, dd))
        self._pend_n += int(du.numel())
        self.max_user_id = max(self.max_user_id, umax)
        self.max_item_id = max(self.max_item_id, imax)
        self._touch()
        if self._pend_n >= _FLUSH_AT:
            self._flush()
        return True

    def _touch(self) -> None:
        self.version += 1
        if self._matrix_cache:
            self._matrix_cache.clear()

    def _seal_scalars(self) -> None:
        if self._pend_scalar:
            a = np.asarray(self._pend_scalar, dtype=np.float64)
            self._pend.append((a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2].copy(), a[:, 3].copy()))
            self._pend_scalar = []

    def _load_host_state(self, keys: np.ndarray, vals: np.ndarray, stamps: np.ndarray, *, max_user_id: int, max_item_id: int,
                         max_timestamp: float, all_item_ids: set, hot_items: Optional[LRUFreqSet] = None) -> None:
        """Replace the store contents with sorted (user << 32 | item, value, stamp) columns held on the host; they are
        uploaded when a GPU is first needed, exactly like an unpickled store.  Used by utils/refpickle.py to take over
        the dict-of-dicts of a model file written by the reference (interactions.py:26)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        assert len(keys) == len(vals) == len(stamps)
        assert len(keys) < 2 or bool((keys[1:] > keys[:-1]).all()), "store keys must be strictly ascending"
        self._keys = self._vals = self._stamps = None
        self._host_state = (keys, np.ascontiguousarray(vals, dtype=np.float64), np.ascontiguousarray(stamps, dtype=np.float64))
        self._n_pairs = len(keys)
        self._pend, self._pend_scalar, self._pend_n, self._pend_upsert = [], [], 0, None
        self.max_user_id, self.max_item_id, self.max_timestamp = int(max_user_id), int(max_item_id), float(max_timestamp)
        self._dev_max_ts = float(max_timestamp)
        self.all_item_ids = set(int(x) for x in all_item_ids)
        if hot_items is not None:
            self.hot_items = hot_items
        self._matrix_cache = {}
        self.version += 1

    def _ensure_device_state(self) -> None:
        if self._host_state is not None:
            k, v, s = self._host_state
            t = D.require_cuda()
            self._keys = D.to_dev(k.view(np.int64))
            self._vals = D.to_dev(v)
            self._stamps = D.to_dev(s)
            self._n_pairs = len(k)
            self._host_state = None

    def _flush(self) -> None:
        """Fold the pending events into the device store (K1)."""
        self._seal_scalars()
        self._ensure_device_state()
        if not self._pend:
            return
        t = D.require_cuda()
        lib = _lib.load()
        cols = []
        for c in range(4):
            parts = [p[c] if not isinstance(p[c], np.ndarray) else D.to_dev(p[c]) for p in self._pend]
            cols.append(parts[0] if len(parts) == 1 else t.cat(parts))
        du, di, dts, dd = cols
        upsert = bool(self._pend_upsert)
        self._pend, self._pend_n, self._pend_upsert = [], 0, None
        n = int(du.numel())
        cap = self._n_pairs + n
        ok = D.empty(cap, t.int64); ov = D.empty(cap, t.float64); os_ = D.empty(cap, t.float64)
        n_out, mts, mu, mi = C.c_int64(0), C.c_double(0), C.c_int32(0), C.c_int32(0)
        rate = self.decay_rate if self.decay_rate is not None else float("nan")
        # carried-in maxima: the device recomputes the same running values the host tracks
        _lib.check(lib.rt_store_fold(D.ptr(du), D.ptr(di), D.ptr(dts), D.ptr(dd), n, int(upsert), float(self.min_value),
                                     float(self.max_value), rate, D.ptr(self._keys), D.ptr(self._vals),
                                     D.ptr(self._stamps), self._n_pairs, float(self._dev_max_ts), 0, 0, D.ptr(ok),
                                     D.ptr(ov), D.ptr(os_), cap, C.byref(n_out), C.byref(mts), C.byref(mu),
                                     C.byref(mi), D.stream_ptr()), "rt_store_fold")
        self._n_pairs = int(n_out.value)
        self._keys, self._vals, self._stamps = ok[:self._n_pairs], ov[:self._n_pairs], os_[:self._n_pairs]
        self._dev_max_ts = float(mts.value)

    # ------------------------------------------------------------------ device matrices
    def device_matrix(self, select_items: Optional[List[int]] = None) -> D.DeviceMatrix:
        """float32 CSR + CSC of the store on the device (K2); cached until the next ingest."""
        self._flush()
        key = None if select_items is None else tuple(sorted(set(int(x) for x in select_items)))
        hit = self._matrix_cache.get(key)
        if hit is not None:
            return hit
        t = D.require_cuda()
        lib = _lib.load()
        n_users, n_items = self.shape
        n = self._n_pairs
        mask = None
        if key is not None:
            m = np.zeros(n_items, dtype=np.uint8)
            sel = np.asarray([x for x in key if 0 <= x < n_items], dtype=np.int64)
            m[sel] = 1
            mask = D.to_dev(m)
        rptr = D.empty(n_users + 1, t.int32); ridx = D.empty(max(n, 1), t.int32); rval = D.empty(max(n, 1), t.float32)
        cptr = D.empty(n_items + 1, t.int32); cidx = D.empty(max(n, 1), t.int32); cval = D.empty(max(n, 1), t.float32)
        ccol = D.empty(max(n, 1), t.int32)
        nnz, nonneg = C.c_int64(0), C.c_int(0)
        rate = self.decay_rate if self.decay_rate is not None else float("nan")
        _lib.check(lib.rt_store_build(D.ptr(self._keys), D.ptr(self._vals), D.ptr(self._stamps), n, rate,
                                      float(self.max_timestamp), n_users, n_items, D.ptr(mask), D.ptr(rptr), D.ptr(ridx),
                                      D.ptr(rval), D.ptr(cptr), D.ptr(cidx), D.ptr(cval), D.ptr(ccol), C.byref(nnz),
                                      C.byref(nonneg), D.stream_ptr()), "rt_store_build")
        X = D.DeviceMatrix(n_users, n_items, int(nnz.value), rptr, ridx, rval, cptr, cidx, cval, ccol, bool(nonneg.value))
        self._matrix_cache[key] = X
        return X

    # ------------------------------------------------------------------ point queries
    def _lookup(self, user_id: int, item_id: int) -> Optional[Tuple[float, float]]:
        self._flush()
        if self._n_pairs == 0 or user_id < 0 or item_id < 0:
            return None
        t = D.require_cuda()
        q = D.to_dev(np.array([(int(user_id) << 32) | int(item_id)], dtype=np.int64))
        v = D.empty(1, t.float64); s = D.empty(1, t.float64); f = D.empty(1, t.uint8)
        _lib.check(_lib.load().rt_store_lookup(D.ptr(self._keys), D.ptr(self._vals), D.ptr(self._stamps), self._n_pairs,
                                               D.ptr(q), 1, D.ptr(v), D.ptr(s), D.ptr(f), D.stream_ptr()),
                   "rt_store_lookup")
        if int(f.item()) == 0:
            return None
        return float(v.item()), float(s.item())

    def has_interaction(self, user_id: int, item_id: int) -> bool:
        return self._lookup(user_id, item_id) is not None

    def get_user_item_rating(self, user_id: int, item_id: int, default_rating: float = 0.0) -> float:
        found = self._lookup(user_id, item_id)
        current, last_ts = found if found is not None else (default_rating, 0.0)
        if current == default_rating:
            return default_rating
        return self._apply_decay(current, last_ts)

    def _user_slice(self, user_id: int):
        """(items int64, vals, stamps) numpy arrays of one user's pairs, ascending item id."""
        self._flush()
        if self._n_pairs == 0 or user_id < 0:
            z = np.zeros(0)
            return z.astype(np.int64), z, z
        t = D.require_cuda()
        bounds = t.tensor([int(user_id) << 32, (int(user_id) + 1) << 32], dtype=t.int64, device=self._keys.device)
        lo, hi = t.searchsorted(self._keys, bounds).tolist()
        k = self._keys[lo:hi].cpu().numpy()
        return k & 0xFFFFFFFF, self._vals[lo:hi].cpu().numpy(), self._stamps[lo:hi].cpu().numpy()

    def get_user_items(self, user_id: int, n_recent: Optional[int] = None) -> List[int]:
        items, _, stamps = self._user_slice(user_id)
        if len(items) == 0:
            return []
        if n_recent is not None and self._n_users_seen() > n_recent:  # condition kept as in interactions.py:168
            order = np.argsort(-stamps, kind="stable")
            return items[order][:n_recent].tolist()
        return items.tolist()

    def _n_users_seen(self) -> int:
        """Distinct users in the store (cached per store version: ``get_user_items(n_recent=...)`` asks per lookup)."""
        hit = getattr(self, "_n_users_cache", None)
        if hit is not None and hit[0] == self.version:
            return hit[1]
        n = len(self.get_all_users())
        self._n_users_cache = (self.version, n)
        return n

    def get_all_item_ids(self) -> List[int]:
        return list(self.all_item_ids)

    def get_all_users(self) -> List[int]:
        self._flush()
        if self._n_pairs == 0:
            return []
        t = D.require_cuda()
        return t.unique(self._keys >> 32).cpu().tolist()

    def get_all_non_interacted_items(self, user_id: int) -> List[int]:
        interacted = self.get_user_items(user_id)
        if len(interacted) == 0:
            return list(self.all_item_ids)
        return list(self.all_item_ids.difference(interacted))

    def get_all_non_negative_items(self, user_id: int) -> List[int]:
        items, vals, stamps = self._user_slice(user_id)
        rated = {int(i): self._apply_decay(float(v), float(s)) if v != 0.0 else 0.0
                 for i, v, s in zip(items, vals, stamps)}
        return [i for i in self.all_item_ids if rated.get(i, 0.0) >= 0.0]

    def get_hot_items(self, n: Optional[int] = None, user_id: Optional[int] = None, filter_interacted: bool = True) -> List[int]:
        interacted: List[int] = []
        if filter_interacted:
            assert user_id is not None, "User ID must be provided to filter interacted items."
            interacted = self.get_user_items(user_id)
        return list(self.hot_items.get_freq_items(n, exclude_items=interacted))

    def get_users_by_items(self, item_ids: List[int]) -> List[int]:
        if not item_ids:
            return []
        X = self.device_matrix()
        t = D.torch()
        cptr = X.cptr.cpu().numpy()
        parts = []
        for it in set(int(x) for x in item_ids):
            if 0 <= it < X.n_items and cptr[it + 1] > cptr[it]:
                parts.append(X.cidx[int(cptr[it]):int(cptr[it + 1])])
        if not parts:
            return []
        return t.unique(t.cat(parts)).cpu().tolist()

    # ------------------------------------------------------------------ host mirrors (scipy)
    def to_csr(self, select_users: Optional[List[int]] = None, include_weights: bool = True):
        import scipy.sparse as sp
        M = self.device_matrix().to_scipy_csr()
        if select_users:
            # the reference appends a user's row once per occurrence and scipy sums duplicates
            # (interactions.py:264-276), so a user listed m times gets m-fold weights
            uniq, counts = np.unique(np.asarray(select_users, dtype=np.int64), return_counts=True)
            ok = (uniq >= 0) & (uniq < M.shape[0])
            mult = np.zeros(M.shape[0], dtype=np.float32)
            mult[uniq[ok]] = counts[ok]
            rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
            sel = mult[rows] > 0
            M = sp.csr_matrix((M.data[sel] * mult[rows[sel]], (rows[sel], M.indices[sel])), shape=M.shape,
                              dtype=np.float32)
        if not include_weights:
            return sp.csr_matrix((np.ones(M.nnz, dtype="int32"), M.indices, M.indptr), shape=M.shape)
        return M

    def to_csc(self, select_items: Optional[List[int]] = None):
        return self.device_matrix(select_items).to_scipy_csc()

    def to_coo(self, select_users: Optional[List[int]] = None, select_items: Optional[List[int]] = None):
        import scipy.sparse as sp
        M = sp.coo_matrix(self.device_matrix(select_items).to_scipy_csr())
        if select_users is not None:
            keep = np.isin(M.row, np.asarray(list(select_users), dtype=np.int64))
            M = sp.coo_matrix((M.data[keep], (M.row[keep], M.col[keep])), shape=M.shape, dtype=np.float32)
        return M

    @property
    def shape(self) -> tuple[int, int]:
        return self.max_user_id + 1, self.max_item_id + 1

    # ------------------------------------------------------------------ pickling
    def __getstate__(self):
        st = dict(self.__dict__)
        if self._pend or self._pend_scalar:
            self._flush()
            st = dict(self.__dict__)
        if self._keys is not None:
            st["_host_state"] = (self._keys.cpu().numpy().view(np.uint64).copy(), self._vals.cpu().numpy().copy(),
                                 self._stamps.cpu().numpy().copy())
        for k in ("_keys", "_vals", "_stamps"):
            st[k] = None
        st["_n_pairs_host"] = self._n_pairs
        st["_matrix_cache"] = {}
        st["_pend"], st["_pend_scalar"], st["_pend_n"], st["_pend_upsert"] = [], [], 0, None
        return st

    def __setstate__(self, st):
        st.pop("_n_pairs_host", None)
        self.__dict__.update(st)
        if self._host_state is not None:
            self._n_pairs = len(self._host_state[0])
