"""climsim_b200 -- B200-native engine for ClimSim's column-emulator hot path.

Host side: Python/PyTorch plumbing that mirrors the reference's model and ``data_utils`` surface.
Device side: hand-written sm_100a CUDA in ``csrc/`` behind the C ABI of ``include/climsim_b200.h``
(``libclimsim_b200.so``).  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from . import _lib  # noqa: F401
from .engine import CNNEngine, MLPEngine  # noqa: F401
from .stream import NpyColumnStream, ResidentColumnStream, StreamPlan  # noqa: F401

__version__ = "0.1.0"
