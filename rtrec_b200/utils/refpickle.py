"""Loading model files written by the reference implementation.

``rtrec.models.SLIM.save`` (/root/reference/rtrec/models/base.py:376-384, slim.py:117-131) pickles a dict of live
reference objects -- ``rtrec.utils.interactions.UserItemInteractions`` (dict-of-dicts store),
``rtrec.models.internal.slim_elastic.SLIMElastic`` (scipy CSC ``item_similarity`` + an sklearn estimator),
``rtrec.utils.identifiers.Identifier`` (x2) and ``rtrec.utils.features.FeatureStore`` -- by class path.  The
reference package is not a dependency of this one, so the unpickler here resolves every ``rtrec.*`` class to a
state-capturing stand-in and ``convert`` rebuilds the equivalent objects of this package from the captured
attribute dicts: the store becomes the sorted (key, value, stamp) columns the device store is built from, W
goes to the device on first use.  Files written by this package load through the same unpickler unchanged.
"""
from __future__ import annotations

import io
import pickle
from collections import OrderedDict
from typing import Any, Dict

import numpy as np


class RefObject:
    """Stand-in for an instance of a reference class: keeps the pickled attribute dict."""
    _ref_path = ""

    def __class_getitem__(cls, item):   # instances of ``IndexedSet[str]`` pickle their ``__orig_class__`` alias
        return cls

    def __setstate__(self, state: Any) -> None:
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict, slots) form
            merged = dict(state[0] or {})
            merged.update(state[1])
            state = merged
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


_shims: Dict[str, type] = {}


def _shim(module: str, name: str) -> type:
    path = f"{module}.{name}"
    if path not in _shims:
        _shims[path] = type(name, (RefObject,), {"_ref_path": path, "__module__": __name__})
    return _shims[path]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):  # noqa: D401
        if module == "rtrec" or module.startswith("rtrec."):
            return _shim(module, name)
        return super().find_class(module, name)


def loads(data: bytes) -> Any:
    """``pickle.loads`` that never imports the reference package."""
    return _Unpickler(io.BytesIO(data)).load()


def is_reference(obj: Any) -> bool:
    return isinstance(obj, RefObject)


# ------------------------------------------------------------------------------------------------ converters
def _identifier(ref: RefObject):
    from .identifiers import Identifier
    out = Identifier(name=ref.__dict__.get("name", "ID"), force_identify=bool(ref.__dict__.get("force_identify", False)))
    out.obj_to_id = dict(ref.__dict__.get("obj_to_id", {}))
    out.id_to_obj = list(ref.__dict__.get("id_to_obj", []))
    out.pass_through = ref.__dict__.get("pass_through", None)
    return out


def _lru(ref: Any, default_capacity: int = 100_000):
    from .lru import LRUFreqSet
    if not is_reference(ref):
        return ref
    out = LRUFreqSet(int(ref.__dict__.get("capacity", default_capacity)))
    out.data = OrderedDict(ref.__dict__.get("data", {}))
    return out


def _feature_store(ref: RefObject):
    """rtrec/utils/features.py:10-15: two IndexedSets (tag vocabularies) + {row id: [tag ids]} maps."""
    from .features import FeatureStore
    out = FeatureStore()
    for side, vocab_attr, map_attr in ((out._users, "user_features", "user_feature_map"),
                                       (out._items, "item_features", "item_feature_map")):
        vocab = ref.__dict__.get(vocab_attr)
        keys = list(vocab.__dict__.get("_index_to_key", [])) if is_reference(vocab) else list(vocab or [])
        side.vocab = {k: n for n, k in enumerate(keys)}
        side.rows = {int(k): [int(x) for x in v] for k, v in dict(ref.__dict__.get(map_attr, {})).items()}
    return out


def _interactions(ref: RefObject):
    """rtrec/utils/interactions.py:15-42: ``interactions`` = {user: {item: (value, timestamp)}} -> the sorted
    (user << 32 | item, value, stamp) columns of the device store.  Values are copied verbatim (they are stored
    un-decayed; decay is applied when a matrix is built, interactions.py:62-79)."""
    from .interactions import UserItemInteractions
    d = ref.__dict__
    out = UserItemInteractions(min_value=d.get("min_value", -5), max_value=d.get("max_value", 10))
    out.decay_rate = d.get("decay_rate", None)
    table = d.get("interactions", {})
    n = sum(len(v) for v in table.values())
    keys = np.empty(n, dtype=np.uint64)
    vals = np.empty(n, dtype=np.float64)
    stamps = np.empty(n, dtype=np.float64)
    p = 0
    for user, row in table.items():
        m = len(row)
        if not m:
            continue
        items = np.fromiter(row.keys(), dtype=np.int64, count=m)
        vs = np.array(list(row.values()), dtype=np.float64).reshape(m, 2)
        keys[p:p + m] = (np.uint64(int(user)) << np.uint64(32)) | items.astype(np.uint64)
        vals[p:p + m] = vs[:, 0]
        stamps[p:p + m] = vs[:, 1]
        p += m
    order = np.argsort(keys[:p], kind="stable")
    out._load_host_state(keys[:p][order], vals[:p][order], stamps[:p][order],
                         max_user_id=int(d.get("max_user_id", 0)), max_item_id=int(d.get("max_item_id", 0)),
                         max_timestamp=float(d.get("max_timestamp", 0.0)), all_item_ids=set(d.get("all_item_ids", ())),
                         hot_items=_lru(d.get("hot_items")))
    return out


def _operator(ref: RefObject):
    """rtrec/models/internal/slim_elastic.py:182-227: configuration attributes + ``item_similarity`` (scipy CSC,
    float64 after a serial ``fit``, float32 otherwise; None before the first fit).  The pickled sklearn estimator
    is dropped: the solver is rebuilt from the configuration."""
    from ..models.internal.slim_elastic import SLIMElastic
    d = ref.__dict__
    cfg = {k: d[a] for k, a in (("optim", "optim_name"), ("eta0", "eta0"), ("alpha", "alpha"), ("l1_ratio", "l1_ratio"),
                                ("positive_only", "positive_only"), ("max_iter", "max_iter"), ("tol", "tol"),
                                ("random_state", "random_state"), ("nn_feature_selection", "nn_feature_selection")) if a in d}
    out = SLIMElastic(cfg)
    W = d.get("item_similarity")
    if W is not None:
        out._set_host_similarity(W)
    return out


def convert(obj: Any) -> Any:
    """Reference stand-in -> object of this package (anything else is returned unchanged)."""
    if not is_reference(obj):
        return obj
    name = obj._ref_path.rsplit(".", 1)[-1]
    if name == "Identifier":
        return _identifier(obj)
    if name == "LRUFreqSet":
        return _lru(obj)
    if name == "FeatureStore":
        return _feature_store(obj)
    if name == "UserItemInteractions":
        return _interactions(obj)
    if name == "SLIMElastic":
        return _operator(obj)
    raise TypeError(f"cannot convert reference object {obj._ref_path}")
