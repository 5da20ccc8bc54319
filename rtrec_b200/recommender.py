"""``Recommender`` -- the DataFrame facade (mirror of /root/reference/rtrec/recommender.py:20-223).

Same methods, arguments, printed lines and return values.  What changes is the data movement:
a DataFrame is handed to the model as four columns (``BaseModel.add_interaction_arrays``) so the
events reach the device store as arrays, and ``evaluate`` asks for the recommendations of many
users per kernel launch instead of 100.
"""
from __future__ import annotations

import math
import time
from typing import Any, Dict, Iterable, Iterator, List, Optional, Tuple

import pandas as pd

from .models.base import BaseModel
from .utils.metrics import compute_scores

_EVAL_DEVICE_BATCH = 16384


class Recommender:
    def __init__(self, model: BaseModel, use_generator: bool = True):
        self.model = model
        self.use_generator = use_generator

    def get_model(self) -> BaseModel:
        return self.model

    def partial_fit(self, user_interactions: Iterable[Tuple[int, int, int, float]], update_interaction: bool = False):
        start_time = time.time()
        self.model.fit(user_interactions, update_interaction=update_interaction, progress_bar=False)
        end_time = time.time()
        print(f"Fit completed in {end_time - start_time:.2f} seconds")
        return self

    def _register_tags(self, user_tags, item_tags) -> None:
        if user_tags:
            for user, tags in user_tags.items():
                self.model.register_user_feature(user, tags)
        if item_tags:
            for item, tags in item_tags.items():
                self.model.register_item_feature(item, tags)

    def _ingest(self, train_data: pd.DataFrame, batch_size: int, update_interaction: bool, record: bool,
                assume_sorted: bool) -> None:
        interaction_df = train_data[["user", "item", "tstamp", "rating"]]
        if not assume_sorted:
            interaction_df = interaction_df.sort_values("tstamp", ascending=True)
        if hasattr(self.model, "add_interaction_arrays"):
            self.model.add_interaction_arrays(interaction_df["user"].to_numpy(), interaction_df["item"].to_numpy(),
                                              interaction_df["tstamp"].to_numpy(), interaction_df["rating"].to_numpy(),
                                              update_interaction=update_interaction, record_interactions=record)
            return
        for batch in Recommender.generate_batches(interaction_df, batch_size, as_generator=self.use_generator):
            self.model.add_interactions(batch, update_interaction=update_interaction, record_interactions=record)

    def fit(self, train_data: pd.DataFrame, user_tags: Optional[Dict[Any, List[str]]] = None,
            item_tags: Optional[Dict[Any, List[str]]] = None, batch_size: int = 1_000, update_interaction: bool = False,
            parallel: bool = False, assume_sorted: bool = True):
        """recommender.py:39-82: ingest, then re-solve the item columns touched by ``train_data``."""
        start_time = time.time()
        self._register_tags(user_tags, item_tags)
        self._ingest(train_data, batch_size, update_interaction, True, assume_sorted)
        self.model._fit_recorded(parallel=parallel, progress_bar=True)
        end_time = time.time()
        print(f"Fit completed in {end_time - start_time:.2f} seconds")
        print(f"Throughput: {len(train_data) / (end_time - start_time):.2f} samples/sec")
        return self

    def bulk_fit(self, train_data: pd.DataFrame, user_tags: Optional[Dict[Any, List[str]]] = None,
                 item_tags: Optional[Dict[Any, List[str]]] = None, batch_size: int = 1_000,
                 update_interaction: bool = False, parallel: bool = True, assume_sorted: bool = True):
        """recommender.py:84-127: ingest, then solve every item column."""
        start_time = time.time()
        self._register_tags(user_tags, item_tags)
        self._ingest(train_data, batch_size, update_interaction, False, assume_sorted)
        self.model.bulk_fit(parallel=parallel, progress_bar=True)
        end_time = time.time()
        print(f"Fit completed in {end_time - start_time:.2f} seconds")
        print(f"Throughput: {len(train_data) / (end_time - start_time):.2f} samples/sec")
        return self

    def recommend(self, user: Any, candidate_items: Optional[List[Any]] = None, user_tags: Optional[List[str]] = None,
                  top_k: int = 10, filter_interacted: bool = True) -> List[Any]:
        return self.model.recommend(user, candidate_items, user_tags, top_k, filter_interacted)

    def recommend_batch(self, users: List[Any], candidate_items: Optional[List[Any]] = None,
                        users_tags: Optional[List[List[str]]] = None, top_k: int = 10,
                        filter_interacted: bool = True) -> List[List[Any]]:
        return self.model.recommend_batch(users, candidate_items, users_tags, top_k, filter_interacted)

    def similar_items(self, query_items: List[Any], query_item_tags: Optional[List[str]] = None, top_k: int = 10,
                      ret_scores: bool = False):
        batch = getattr(self.model, "similar_items_batch", None)
        if batch is not None and query_item_tags is None and len(query_items) > 1:
            return batch(query_items, top_k=top_k, ret_scores=ret_scores)   # one launch for the whole list
        return [self.model.similar_items(item, query_item_tags, top_k, ret_scores) for item in query_items]

    def evaluate(self, test_data: pd.DataFrame, user_tags: Optional[Dict[Any, List[str]]] = None,
                 recommend_size: int = 10, batch_size=100, filter_interacted: bool = True) -> Dict[str, float]:
        """recommender.py:163-200.  Models that offer ``evaluate_device`` (SLIM) get the whole evaluation on the device:
        one scoring launch for all test users and one metrics kernel (``rt_eval_metrics``); the values equal the
        list-based loop below bit for bit.  Otherwise ``batch_size`` is honoured as a lower bound: the scoring kernel is
        fed at least 16,384 users per launch, which returns the same lists as 100 at a time."""
        dev = getattr(self.model, "evaluate_device", None)
        if dev is not None and not user_tags and len(test_data):
            # device path: ground truth as one flat array grouped by user (same grouping and within-user order as the
            # groupby below), top-k lists and the nine metrics stay on the device
            try:
                import numpy as np
                codes, uniques = pd.factorize(test_data["user"], sort=True)
                order = np.argsort(codes, kind="stable")
                gptr = np.concatenate([[0], np.cumsum(np.bincount(codes, minlength=len(uniques)))])
                res = dev(uniques.tolist() if uniques.dtype.kind == "O" else np.asarray(uniques), gptr,
                          test_data["item"].to_numpy()[order], recommend_size, filter_interacted)
            except TypeError:
                res = None   # user keys that cannot be ordered: pandas' groupby decides below
            if res is not None:
                return res
        grouped = test_data.groupby("user")["item"].apply(list).to_dict()
        users = list(grouped.keys())
        step = max(int(batch_size), _EVAL_DEVICE_BATCH)

        def pairs() -> Iterable[Tuple[List[Any], List[Any]]]:
            for i in range(0, len(users), step):
                chunk = users[i:i + step]
                tags = [user_tags.get(u, []) for u in chunk] if user_tags else None
                recs = self.recommend_batch(chunk, users_tags=tags, top_k=recommend_size, filter_interacted=filter_interacted)
                for u, rec in zip(chunk, recs):
                    yield rec, grouped[u]

        return compute_scores(pairs(), recommend_size)

    @staticmethod
    def generate_batches(df: pd.DataFrame, batch_size: int = 1_000, as_generator: bool = False) -> Iterator[Iterable[Tuple[int, int, int, float]]]:
        for start in range(0, len(df), batch_size):
            rows = df.iloc[start:start + batch_size].itertuples(index=False, name=None)
            yield rows if as_generator else list(rows)
