"""Seeded weights / inputs shared by the oracle and the CUDA path (defined once in dvc/synth.py)."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-exemplar-based-video-colorization_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dvc.synth import *  # noqa: F401,F403,E402
from dvc.synth import make_lab, make_state_dict, net_shapes  # noqa: F401,E402
