"""ctypes binding of oracle/_ref/libais_ref.so: the gr-ais reference's OWN block classes
(/root/reference/lib/*_impl.cc compiled unmodified by oracle/ref_build/Makefile against a stub of
the GNU Radio runtime; the GNU Radio / VOLK kernels underneath are the oracle's restatements).

TEST INFRASTRUCTURE ONLY -- used by tests/ to pin oracle/ais_oracle.c and the CUDA path to the
reference's code, and by bench.py's CPU arm.  The product package never imports this module.

The classes mirror oracle.CorrEstBlock / oracle.MskBlock call for call, so a test can run the same
body over both.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libais_ref.so")
REFERENCE_ROOT = os.environ.get("GR_AIS_REFERENCE", "/root/reference")

_fp, _c64, TAG_DTYPE = _o._fp, _o._c64, _o.TAG_DTYPE


def build(force=False):
    """Compile the reference sources (needs REFERENCE_ROOT; this container only)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "lib")):
        raise FileNotFoundError("reference sources not present at %s" % REFERENCE_ROOT)
    _o.build()
    cmd = ["make", "-C", os.path.join(_HERE, "ref_build"), "-s", "REF=" + REFERENCE_ROOT]
    if force:
        cmd.insert(1, "-B")
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def available():
    """True when libais_ref.so exists (prebuilt, or buildable because the reference is here)."""
    if os.path.exists(_LIB_PATH):
        return True
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "lib")):
        try:
            build()
        except Exception:
            return False
        return os.path.exists(_LIB_PATH)
    return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(_LIB_PATH)
        _o.lib()  # libais_oracle.so first (rpath finds it too)
        L = C.CDLL(_LIB_PATH)
        L.ref_blocks.restype = C.c_void_p
        L.ref_corr_est_new.restype = C.c_void_p
        L.ref_corr_est_new.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_uint, C.c_float]
        L.ref_corr_est_delete.argtypes = [C.c_void_p]
        L.ref_corr_est_hints.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ref_corr_est_output_multiple.argtypes = [C.c_void_p]
        L.ref_corr_est_symbols.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_corr_est_set_symbols.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_corr_est_work.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p]
        L.ref_msk_new.restype = C.c_void_p
        L.ref_msk_new.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p]
        L.ref_msk_delete.argtypes = [C.c_void_p]
        for name in ("ref_msk_get_sps", "ref_msk_get_gain", "ref_msk_get_limit"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_msk_relative_rate.restype = C.c_double
        L.ref_msk_relative_rate.argtypes = [C.c_void_p]
        for name in ("ref_msk_set_sps", "ref_msk_set_gain", "ref_msk_set_limit"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_float]
        L.ref_msk_forecast.argtypes = [C.c_void_p, C.c_int]
        L.ref_msk_general_work.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        L.ref_freqest_new.restype = C.c_void_p
        L.ref_freqest_new.argtypes = [C.c_float, C.c_int, C.c_int]
        L.ref_freqest_delete.argtypes = [C.c_void_p]
        L.ref_freqest_work.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_invert_work.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_pdu_to_nmea.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def blocks():
    """ao_blocks pointer for oracle.demod_chain(..., blocks=...) / DemodStream / demod_chain_batch."""
    return C.c_void_p(lib().ref_blocks())


def sources_sha256():
    with open(os.path.join(_HERE, "_ref", "SOURCES.sha256")) as f:
        return f.read()


class CorrEstBlock:
    """gr::ais::corr_est_cc (lib/corr_est_cc_impl.cc), one work() call at a time."""

    def __init__(self, symbols, sps, mark_delay, threshold=0.9):
        symbols = _c64(symbols)
        self._h = lib().ref_corr_est_new(_fp(symbols), len(symbols), C.c_float(sps),
                                         C.c_uint(mark_delay), C.c_float(threshold))
        if not self._h:
            raise RuntimeError("corr_est_cc::make threw")

    def __del__(self):
        try:
            if self._h:
                lib().ref_corr_est_delete(self._h)
                self._h = None
        except Exception:
            pass

    def hints(self):
        v = [C.c_int(0) for _ in range(5)]
        lib().ref_corr_est_hints(self._h, *[C.addressof(x) for x in v])
        return dict(zip(("history", "output_multiple", "max_noutput_items", "sample_delay0",
                         "sample_delay1"), [x.value for x in v]))

    @property
    def L(self):
        return self.hints()["history"] - 1

    @property
    def nsamples(self):
        return int(lib().ref_corr_est_output_multiple(self._h))

    def symbols(self):
        n = lib().ref_corr_est_symbols(self._h, None, 0)
        out = np.zeros(n, dtype=np.complex64)
        lib().ref_corr_est_symbols(self._h, _fp(out), n)
        return out

    def set_symbols(self, symbols):
        symbols = _c64(symbols)
        lib().ref_corr_est_set_symbols(self._h, _fp(symbols), len(symbols))

    def work(self, n, inbuf, nitems_written=0, two_ports=False, max_tags=4096):
        inbuf = _c64(inbuf)
        assert len(inbuf) >= n + self.L
        out0 = np.empty(n, dtype=np.complex64)
        corr = np.empty(n, dtype=np.complex64)
        mag = np.empty(n, dtype=np.float32)
        tags = np.zeros(max_tags, dtype=TAG_DTYPE)
        nt = C.c_int(0)
        rc = lib().ref_corr_est_work(self._h, int(n), _fp(inbuf), C.c_uint64(nitems_written),
                                     _fp(out0), _fp(corr), _fp(mag), int(bool(two_ports)), _fp(tags),
                                     int(max_tags), C.addressof(nt))
        if rc < 0:
            raise ValueError("noutput_items must be a multiple of the output multiple")
        if nt.value > max_tags:
            raise RuntimeError("tag buffer overflow")
        return out0, corr, mag, tags[:nt.value].copy()


class MskBlock:
    """gr::ais::msk_timing_recovery_cc (lib/msk_timing_recovery_cc_impl.cc)."""

    def __init__(self, sps, gain, limit, osps=1):
        st = C.c_int(0)
        self._h = lib().ref_msk_new(C.c_float(sps), C.c_float(gain), C.c_float(limit), int(osps),
                                    C.addressof(st))
        if st.value == -1:
            raise IndexError("Gain must be positive")  # std::out_of_range
        if st.value == -2:
            raise IndexError("osps must be 1 or 2")
        if not self._h:
            raise RuntimeError("msk_timing_recovery_cc::make threw")

    def __del__(self):
        try:
            if self._h:
                lib().ref_msk_delete(self._h)
                self._h = None
        except Exception:
            pass

    def get_sps(self):
        return float(lib().ref_msk_get_sps(self._h))

    def get_gain(self):
        return float(lib().ref_msk_get_gain(self._h))

    def get_limit(self):
        return float(lib().ref_msk_get_limit(self._h))

    def set_sps(self, v):
        lib().ref_msk_set_sps(self._h, C.c_float(v))

    def set_limit(self, v):
        lib().ref_msk_set_limit(self._h, C.c_float(v))

    def set_gain(self, v):
        if lib().ref_msk_set_gain(self._h, C.c_float(v)):
            raise IndexError("Gain must be positive")

    def relative_rate(self):
        return float(lib().ref_msk_relative_rate(self._h))

    def forecast(self, noutput_items):
        return int(lib().ref_msk_forecast(self._h, int(noutput_items)))

    def general_work(self, noutput_items, inbuf, tags=None, nitems_read=0):
        inbuf = _c64(inbuf)
        if tags is None:
            tags = np.zeros(0, dtype=TAG_DTYPE)
        tags = np.ascontiguousarray(tags, dtype=TAG_DTYPE)
        out = np.empty(max(noutput_items, 1), dtype=np.complex64)
        err = np.empty(max(noutput_items, 1), dtype=np.float32)
        mu = np.empty(max(noutput_items, 1), dtype=np.float32)
        consumed = C.c_int(0)
        k = lib().ref_msk_general_work(self._h, int(noutput_items), len(inbuf), _fp(inbuf),
                                       C.c_uint64(nitems_read), _fp(tags), len(tags), _fp(out),
                                       _fp(err), _fp(mu), C.addressof(consumed))
        if k < 0:
            raise RuntimeError("mmse interpolator index out of range")
        return out[:k].copy(), err[:k].copy(), mu[:k].copy(), consumed.value


def freqest_work(spec, sample_rate=48000.0, data_rate=9600, fftlen=1024):
    """gr::ais::freqest::work over [nvec, fftlen] complex64 -> hz[nvec] float32"""
    spec = _c64(spec).reshape(-1, fftlen)
    h = lib().ref_freqest_new(C.c_float(sample_rate), int(data_rate), int(fftlen))
    out = np.zeros(len(spec), dtype=np.float32)
    try:
        lib().ref_freqest_work(h, _fp(spec), len(spec), _fp(out))
    finally:
        lib().ref_freqest_delete(h)
    return out


def invert(b):
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(len(b), dtype=np.uint8)
    lib().ref_invert_work(_fp(b), len(b), _fp(out))
    return out


def pdu_to_nmea(data: bytes, designator="A") -> str:
    buf = C.create_string_buffer(8192)
    n = lib().ref_pdu_to_nmea(designator.encode(), bytes(data), len(data), buf, 8192)
    if n < 0:
        raise ValueError("pdu_to_nmea harness: %d" % n)
    return buf.raw[:n].decode("latin-1")
