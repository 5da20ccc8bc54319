"""``BaseModel`` -- the plugin base ``Recommender`` drives (host-side Python).

Behavioural mirror of /root/reference/rtrec/models/base.py:21-423: id resolution, the cold/hot
user split, candidate filtering, per-event error swallowing and pickle save/load are bookkeeping
that must stay bit-exact, so they stay on the host.  The only addition is
``add_interaction_arrays``: the column-oriented ingest the DataFrame facade uses so that a batch
reaches the device store as arrays instead of one Python call per event.
"""
from __future__ import annotations

import logging
import pickle
from abc import ABC, abstractmethod
from io import BytesIO
from typing import Any, Iterable, List, Optional, Tuple, Union

import numpy as np

from ..utils.features import FeatureStore
from ..utils.identifiers import Identifier
from ..utils.interactions import UserItemInteractions

FileLike = Union[BytesIO, Any]
_INT32_MAX = 2**31 - 1


class BaseModel(ABC):
    def __init__(self, **kwargs: Any):
        self.interactions = UserItemInteractions(**kwargs)
        self.user_ids = Identifier(**kwargs)
        self.item_ids = Identifier(**kwargs)
        self.feature_store = FeatureStore()

    # ------------------------------------------------------------------ tags (not on the SLIM path)
    def register_user_feature(self, user: Any, user_tags: List[str]) -> int:
        user_id = self.user_ids.identify(user)
        self.feature_store.put_user_features(user_id, user_tags)
        return user_id

    def clear_user_features(self, user_ids: Optional[List[int]] = None) -> None:
        self.feature_store.clear_user_features(user_ids)

    def register_item_feature(self, item: Any, item_tags: List[str]) -> int:
        item_id = self.item_ids.identify(item)
        self.feature_store.put_item_features(item_id, item_tags)
        return item_id

    def clear_item_features(self, item_ids: Optional[List[int]] = None) -> None:
        self.feature_store.clear_item_features(item_ids)

    # ------------------------------------------------------------------ ingest
    def add_interactions(self, interactions: Iterable[Tuple[Any, Any, float, float]], update_interaction: bool = False,
                         record_interactions: bool = False) -> None:
        """base.py:72-94: malformed events are skipped with a warning, the rest are stored."""
        users, items, stamps, ratings = [], [], [], []
        for event in interactions:
            try:
                user, item, tstamp, rating = event
                user_id = self.user_ids.identify(user)
                item_id = self.item_ids.identify(item)
                tstamp, rating = float(tstamp), float(rating)
                # the device store indexes with int32: an id outside [0, 2^31) is this event's error, not the batch's
                if not (0 <= user_id <= _INT32_MAX and 0 <= item_id <= _INT32_MAX):
                    raise ValueError(f"id outside [0, 2^31): ({user_id}, {item_id})")
            except Exception as e:  # noqa: BLE001 - same breadth as the reference
                logging.warning(f"Error processing interaction: {e}")
                continue
            users.append(user_id); items.append(item_id); stamps.append(tstamp); ratings.append(rating)
        if not users:
            return
        try:
            self.interactions.add_interactions_batch(users, items, stamps, ratings, upsert=update_interaction)
        except Exception as e:  # noqa: BLE001
            logging.warning(f"Error processing interaction: {e}")
            return
        if record_interactions:
            for user_id, item_id, tstamp, rating in zip(users, items, stamps, ratings):
                self._record_interactions(user_id, item_id, tstamp, rating)

    def add_interaction_arrays(self, users, items, tstamps, ratings, update_interaction: bool = False,
                               record_interactions: bool = False) -> None:
        """Column-oriented ``add_interactions`` (same ids, same store state, same recorded items)."""
        try:
            user_ids = self.user_ids.identify_many(users)
            item_ids = self.item_ids.identify_many(items)
            # (the store validates the id ranges itself -- on the device for large batches -- before it changes state)
            self.interactions.add_interactions_batch(user_ids, item_ids, np.asarray(tstamps, dtype=np.float64),
                                                     np.asarray(ratings, dtype=np.float64), upsert=update_interaction)
        except (ValueError, TypeError) as e:
            # something in the columns is malformed (mixed id kinds, an id outside [0, 2^31), a non-numeric rating):
            # replay them event by event, so that only the offending events are skipped with a warning (base.py:85-94)
            logging.warning(f"Column ingest fell back to the per-event path: {e}")
            self.add_interactions(zip(list(users), list(items), list(tstamps), list(ratings)),
                                  update_interaction=update_interaction, record_interactions=record_interactions)
            return
        if record_interactions:
            self._record_interaction_arrays(user_ids, item_ids)

    def _record_interaction_arrays(self, user_ids: np.ndarray, item_ids: np.ndarray) -> None:
        for u, i in zip(user_ids.tolist(), item_ids.tolist()):
            self._record_interactions(u, i, 0.0, 0.0)

    @abstractmethod
    def _record_interactions(self, user_id: int, item_id: int, tstamp: float, rating: float) -> None:
        raise NotImplementedError("_record_interactions method must be implemented in the derived class")

    def fit(self, interactions: Iterable[Tuple[Any, Any, float, float]], update_interaction: bool = False,
            progress_bar: bool = True):
        self.add_interactions(interactions, update_interaction=update_interaction, record_interactions=True)
        return self._fit_recorded(progress_bar=progress_bar)

    @abstractmethod
    def _fit_recorded(self, parallel: bool = False, progress_bar: bool = True):
        raise NotImplementedError("_fit_recorded method must be implemented in the derived class")

    @abstractmethod
    def bulk_fit(self, parallel: bool = True, progress_bar: bool = True):
        raise NotImplementedError("bulk_fit method must be implemented in the derived class")

    # ------------------------------------------------------------------ recommend
    def _known_candidates(self, candidate_items: Optional[List[Any]]) -> Optional[List[int]]:
        if candidate_items is None:
            return None
        out = []
        for item in candidate_items:
            item_id = self.item_ids.get_id(item)
            if item_id is None:
                continue
            if self.item_ids.pass_through and not (0 <= item_id <= self.interactions.max_item_id):
                continue  # unseen (or negative) integer id: not a candidate
            out.append(item_id)
        return out or None

    def _resolve_user(self, user: Any) -> Optional[int]:
        uid = self.user_ids.get_id(user)
        if uid is not None and self.user_ids.pass_through and not (0 <= uid <= self.interactions.max_user_id):
            return None  # unseen (or negative) integer id: cold start
        return uid

    def recommend(self, user: Any, candidate_items: Optional[List[Any]] = None, user_tags: Optional[List[str]] = None,
                  top_k: int = 10, filter_interacted: bool = True) -> List[Any]:
        """base.py:135-173."""
        candidate_item_ids = self._known_candidates(candidate_items)
        user_id = self._resolve_user(user)
        if user_id is None:
            hot = self.interactions.get_hot_items(top_k, filter_interacted=False)
            if candidate_item_ids is not None:
                hot = [i for i in hot if i in candidate_item_ids]
            return hot
        rec = self._recommend(user_id, candidate_item_ids=candidate_item_ids, user_tags=user_tags, top_k=top_k,
                              filter_interacted=filter_interacted)
        return [self.item_ids.get(i) for i in rec]

    @abstractmethod
    def _recommend(self, user_id: int, candidate_item_ids: Optional[List[int]] = None, user_tags: Optional[List[str]] = None,
                   top_k: int = 10, filter_interacted: bool = True) -> List[int]:
        raise NotImplementedError("_recommend method must be implemented in the derived class")

    def recommend_batch(self, users: List[Any], candidate_items: Optional[List[Any]] = None,
                        users_tags: Optional[List[List[str]]] = None, top_k: int = 10,
                        filter_interacted: bool = True) -> List[List[Any]]:
        """base.py:188-269."""
        cold_slots, hot_slots, hot_ids = [], [], []
        cold_ids: List[Optional[int]] = []
        if self.user_ids.pass_through and len(users) >= 1024:
            # integer pass-through ids (identifiers.py:69-72): resolve the whole list at once
            arr = np.asarray(users)
            if arr.dtype.kind in "iu" and arr.ndim == 1 and (arr.min() >= 0) and (arr.max() <= self.interactions.max_user_id):
                hot_ids = arr
                users = ()
        for slot, user in enumerate(users):
            uid = self._resolve_user(user)
            if uid is None:
                cold_ids.append(self.handle_unknown_user(user))
                cold_slots.append(slot)
            else:
                hot_ids.append(uid)
                hot_slots.append(slot)
        candidate_item_ids = self._known_candidates(candidate_items)
        to_items = self.item_ids.get
        if not cold_slots:
            batch = self._recommend_hot_batch(hot_ids, candidate_item_ids=candidate_item_ids, users_tags=users_tags,
                                              top_k=top_k, filter_interacted=filter_interacted)
            if self.item_ids.pass_through:
                return batch  # integer ids map to themselves (identifiers.py:69-72)
            return [[to_items(i) for i in row] for row in batch]
        results: List[List[Any]] = [[] for _ in users]
        cold_tags = [users_tags[s] for s in cold_slots] if users_tags else None
        cold = self._recommend_cold_batch(cold_ids, candidate_item_ids=candidate_item_ids, users_tags=cold_tags, top_k=top_k)
        for row, slot in zip(cold, cold_slots):
            results[slot] = [to_items(i) for i in row]
        if hot_slots:
            hot_tags = [users_tags[s] for s in hot_slots] if users_tags else None
            hot = self._recommend_hot_batch(hot_ids, candidate_item_ids=candidate_item_ids, users_tags=hot_tags,
                                            top_k=top_k, filter_interacted=filter_interacted)
            for row, slot in zip(hot, hot_slots):
                results[slot] = [to_items(i) for i in row]
        return results

    def handle_unknown_user(self, user: Any) -> Optional[int]:
        return None

    def _recommend_cold_batch(self, user_ids: List[Optional[int]], candidate_item_ids: Optional[List[int]] = None,
                              users_tags: Optional[List[List[str]]] = None, top_k: int = 10) -> List[List[int]]:
        hot = self.interactions.get_hot_items(top_k, filter_interacted=False)
        if candidate_item_ids is not None:
            hot = [i for i in hot if i in candidate_item_ids]
        return [hot for _ in user_ids]

    def _recommend_hot_batch(self, user_ids: List[int], candidate_item_ids: Optional[List[int]] = None,
                             users_tags: Optional[List[List[str]]] = None, top_k: int = 10,
                             filter_interacted: bool = True) -> List[List[int]]:
        if users_tags:
            assert len(user_ids) == len(users_tags), f"Number of user tags must match the number of users. Got {len(user_ids)} users and {len(users_tags)} user tags."
            return [self._recommend(u, candidate_item_ids=candidate_item_ids, user_tags=tags, top_k=top_k,
                                    filter_interacted=filter_interacted) for u, tags in zip(user_ids, users_tags)]
        return [self._recommend(u, candidate_item_ids=candidate_item_ids, top_k=top_k, filter_interacted=filter_interacted)
                for u in user_ids]

    # ------------------------------------------------------------------ similar items / reverse lookup
    def similar_items(self, query_item: Any, query_item_tags: Optional[List[str]] = None, top_k: int = 10,
                      ret_scores: bool = False):
        """base.py:320-340 (note: ``identify`` registers an unknown query item, as the reference does)."""
        query_item_id = self.item_ids.identify(query_item)
        if query_item_id is None:
            return []
        pairs = self._similar_items(query_item_id, query_item_tags=query_item_tags, top_k=top_k)
        if ret_scores:
            return [(self.item_ids.get(i), s) for i, s in pairs]
        return [self.item_ids.get(i) for i, _ in pairs]

    def get_users_by_items(self, items: List[Any]) -> List[Any]:
        item_ids = [i for i in (self.item_ids.get_id(it) for it in items) if i is not None]
        if not item_ids:
            return []
        return [self.user_ids.get(u) for u in self.interactions.get_users_by_items(item_ids)]

    @abstractmethod
    def _similar_items(self, query_item_id: int, query_item_tags: Optional[List[str]] = None, top_k: int = 10) -> List[Tuple[int, float]]:
        raise NotImplementedError("_similar_items method must be implemented in the derived class")

    # ------------------------------------------------------------------ persistence (base.py:376-405)
    def save(self, f: FileLike) -> int:
        return f.write(pickle.dumps(self._serialize(), protocol=pickle.HIGHEST_PROTOCOL))

    @classmethod
    def load(cls, f: FileLike):
        """Reads a model file written by ``save`` -- of this package or of the reference implementation
        (rtrec.models.*.save): reference objects are rebuilt as this package's (utils/refpickle.py)."""
        from ..utils import refpickle
        data = refpickle.loads(f.read())
        if isinstance(data, dict):
            data = {k: refpickle.convert(v) for k, v in data.items()}
        return cls._deserialize(data)

    @classmethod
    def loads(cls, data: bytes):
        return cls.load(BytesIO(data))

    @abstractmethod
    def _serialize(self) -> dict:
        raise NotImplementedError("_serialize method must be implemented in the derived class")

    @classmethod
    @abstractmethod
    def _deserialize(cls, data: dict):
        raise NotImplementedError("_deserialize method must be implemented in the derived class")
