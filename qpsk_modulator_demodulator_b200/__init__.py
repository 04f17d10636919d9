"""qpsk_modulator_demodulator_b200 — B200-native hot path of the C# QPSK modem.

csrc/   hand-written CUDA (sm_100a) kernels + the C ABI of include/qpskcuda.h -> lib/libqpskcuda.so
api.py / modem.py  host-side mirror of the reference's class surface over that C ABI (ctypes)

(The directory uses underscores so that it is an importable Python package; the project name is
qpsk-modulator-demodulator_b200.)
"""
from .api import *  # noqa: F401,F403
from .modem import *  # noqa: F401,F403
from . import api, modem, _native  # noqa: F401
