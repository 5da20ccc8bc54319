"""Object <-> dense id mapping (behavioural mirror of /root/reference/rtrec/utils/identifiers.py:6-90).

Integers pass through unchanged unless ``force_identify`` is set; any other hashable gets the next
dense id on first sight.  Mixing the two kinds raises ``ValueError("Mixed types detected ...")``.
``identify_many`` is the vectorised form used by the batched ingest path and yields exactly the
ids a loop of ``identify`` calls would.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np


class IdentifierError(Exception):
    def __init__(self, id_name: str, obj_id: int):
        super().__init__(f"Identifier not found for {id_name}: {obj_id}")


def _is_int(obj: Any) -> bool:
    return isinstance(obj, (int, np.integer))


class Identifier:
    def __init__(self, name: str = "ID", force_identify: bool = False, **kwargs: Any) -> None:
        self.name = name
        self.force_identify = force_identify
        self.obj_to_id: dict[Any, int] = {}
        self.id_to_obj: list[Any] = []
        # None = undecided, True = integer pass-through, False = dictionary ids
        self.pass_through: Optional[bool] = False if force_identify else None

    def _mixed(self, obj: Any) -> ValueError:
        return ValueError(f"Mixed types detected for {self.name}: {obj}")

    def identify(self, obj: Any) -> int:
        if not self.force_identify and _is_int(obj):
            if self.pass_through is False:
                raise self._mixed(obj)
            self.pass_through = True
            return int(obj)
        if self.pass_through is True:
            raise self._mixed(obj)
        known = self.obj_to_id.get(obj)
        if known is not None:
            return known
        new_id = len(self.id_to_obj)
        self.obj_to_id[obj] = new_id
        self.id_to_obj.append(obj)
        self.pass_through = False
        return new_id

    def identify_many(self, objs) -> np.ndarray:
        """ids of a whole column, in order; equals ``[self.identify(o) for o in objs]``."""
        if isinstance(objs, np.ndarray):
            arr = objs
        else:
            objs = list(objs)
            # a list of Python objects keeps its element types (np.asarray would silently turn [1, "a"] into strings and
            # hide the "Mixed types" error the per-event path raises)
            arr = np.asarray(objs) if all(_is_int(o) for o in objs[:64]) else np.empty(0, dtype=object)
            if arr.dtype.kind not in "iu":
                arr = np.empty(len(objs), dtype=object)
                arr[:] = objs
        if arr.dtype.kind in "iu" and not self.force_identify:
            if self.pass_through is False:
                raise self._mixed(arr[0] if len(arr) else None)
            if len(arr):
                self.pass_through = True
            return arr.astype(np.int64, copy=False)
        out = np.empty(len(arr), dtype=np.int64)
        ident = self.identify
        for k, o in enumerate(arr.tolist() if arr.dtype.kind != "O" else arr):
            out[k] = ident(o)
        return out

    def get_id(self, obj: Any) -> Optional[int]:
        if not self.force_identify and _is_int(obj):
            if not self.pass_through:
                raise self._mixed(obj)
            return int(obj)
        return self.obj_to_id.get(obj)

    def get(self, obj_id: int) -> Any:
        if self.pass_through:
            return obj_id
        if 0 <= obj_id < len(self.id_to_obj):
            return self.id_to_obj[obj_id]
        raise IdentifierError(self.name, obj_id)

    def get_or_default(self, obj_id: int, default: Optional[Any] = None) -> Any:
        if self.pass_through:
            return obj_id
        if 0 <= obj_id < len(self.id_to_obj):
            return self.id_to_obj[obj_id]
        return default

    def __getitem__(self, obj_id: int) -> Any:
        return self.get(obj_id)
