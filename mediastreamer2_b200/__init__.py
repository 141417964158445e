"""mediastreamer2_b200 — B200-native batched implementation of mediastreamer2's per-tick DSP hot path.

The product is ``lib/libmsb200dsp.so`` (C ABI in ``include/msb200dsp.h``, CUDA sources in ``csrc/``) plus the
MSFilterDesc plugin in ``plugin/``. This Python package is test/bench plumbing over that ABI.
"""
__version__ = "0.1.0"
