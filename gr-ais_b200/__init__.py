"""gr-ais_b200: B200-native (sm_100a) implementation of the gr-ais IQ-demod hot path.

Only what the path needs lives here: csrc/ (CUDA kernels + the C-ABI library
libb200ais.so), the ctypes binding of that C-ABI, the host-side mirrors of the
reference's operator interface (blocks.py, ais_demod.py) and a synthetic-traffic
generator for tests and benchmarks (synth.py).
"""
__version__ = "0.1.0"
