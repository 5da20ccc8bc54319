"""rtrec_b200 -- B200-native SLIM hot path behind the rtrec API.

    from rtrec_b200.models import SLIM
    from rtrec_b200.recommender import Recommender

Importing never touches the GPU; the first compute call loads ``csrc/librtrec_b200.so`` and
raises ``RtrecB200Error`` if the library or a CUDA device is missing (there is no CPU fallback).
"""
from ._lib import RtrecB200Error  # noqa: F401

__version__ = "0.1.0"
