"""Synthetic workload generator: moved to ``rtrec_b200/utils/synth.py`` (it is input data for the bench, the tools and
the tests, not part of the checker); this module keeps the old import path for the tests."""
from rtrec_b200.utils.synth import *  # noqa: F401,F403
from rtrec_b200.utils.synth import SHAPES, synth_events, synth_shape  # noqa: F401
