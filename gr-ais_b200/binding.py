"""ctypes binding of the C-ABI in include/b200ais.h (libb200ais.so).

Fails loudly when the CUDA library is missing or a call reports an error: there is
no CPU fallback anywhere in this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# B200AIS_LIB: another build of the same library (kernel experiments); the default is the in-tree one
LIB_PATH = os.environ.get("B200AIS_LIB") or os.path.join(_PKG, "libb200ais.so")
CSRC = os.path.join(_PKG, "csrc")

OK = 0
E_INVALID, E_RANGE, E_CUDA, E_NOMEM, E_TAG_OVERFLOW, E_INTERP, E_OUT_OVERFLOW = -1, -2, -3, -4, -5, -6, -7
TAG_CORR_START, TAG_PHASE_EST, TAG_TIME_EST, TAG_CORR_EST = 0, 1, 2, 3
STAGE_FREQSYNC, STAGE_AGC = 1, 2
TAP_FHAT, TAP_AGC, TAP_SYM, TAP_ERR, TAP_MU, TAP_SOFT, TAP_MASK = range(7)

TAG_DTYPE = np.dtype([("offset", "<u8"), ("key", "<i4"), ("port", "<i4"), ("value", "<f8")])

# every symbol include/b200ais.h declares (tests check the library exports all of them)
EXPORTS = [
    "b200ais_version", "b200ais_last_error", "b200ais_device_count", "b200ais_set_device",
    "b200ais_host_alloc", "b200ais_host_free", "b200ais_launch_count",
    "b200ais_corr_est_create", "b200ais_corr_est_destroy", "b200ais_corr_est_set_symbols",
    "b200ais_corr_est_symbols", "b200ais_corr_est_output_multiple", "b200ais_corr_est_history",
    "b200ais_corr_est_mark_delay", "b200ais_corr_est_threshold", "b200ais_corr_est_work",
    "b200ais_corr_est_work_dev",
    "b200ais_msk_create", "b200ais_msk_destroy", "b200ais_msk_set_gain", "b200ais_msk_get_gain",
    "b200ais_msk_set_limit", "b200ais_msk_get_limit", "b200ais_msk_set_sps", "b200ais_msk_get_sps",
    "b200ais_msk_forecast", "b200ais_msk_reset", "b200ais_msk_general_work",
    "b200ais_msk_general_work_dev",
    "b200ais_freqest_create", "b200ais_freqest_destroy", "b200ais_freqest_work",
    "b200ais_freqest_work_dev",
    "b200ais_invert_work", "b200ais_invert_work_dev",
    "b200ais_demod_default_config", "b200ais_demod_create", "b200ais_demod_destroy",
    "b200ais_demod_max_bits", "b200ais_demod_work", "b200ais_demod_work_sc16", "b200ais_demod_work_dev",
    "b200ais_demod_status", "b200ais_demod_enable_taps", "b200ais_demod_tap",
    "b200ais_demod_read_tap", "b200ais_demod_profile", "b200ais_demod_stage_ms",
    "b200ais_demod_set_overlap", "b200ais_demod_stream_reset", "b200ais_demod_stream_max_bits",
    "b200ais_demod_stream_work", "b200ais_demod_stream_work_dev", "b200ais_demod_stream_pending",
    "b200ais_demod_enqueue_dev", "b200ais_demod_join", "b200ais_demod_set_symbols",
    "b200ais_firdes_low_pass", "b200ais_xlat_create", "b200ais_xlat_destroy",
    "b200ais_xlat_history", "b200ais_xlat_decimation", "b200ais_xlat_set_center_freq",
    "b200ais_xlat_set_taps", "b200ais_xlat_reset", "b200ais_xlat_work", "b200ais_xlat_work_dev",
    "b200ais_hdlc_create", "b200ais_hdlc_destroy", "b200ais_hdlc_reset", "b200ais_hdlc_work",
    "b200ais_hdlc_work_dev", "b200ais_hdlc_status",
    "b200ais_nmea_slot_bytes", "b200ais_nmea_format", "b200ais_nmea_format_dev",
    "b200ais_demod_stream_stage", "b200ais_demod_stream_work_staged",
    "b200ais_rx_default_config", "b200ais_rx_create", "b200ais_rx_destroy", "b200ais_rx_reset",
    "b200ais_rx_decimation", "b200ais_rx_channels", "b200ais_rx_samples_per_symbol",
    "b200ais_rx_sentence_slot", "b200ais_rx_work", "b200ais_rx_work_dev", "b200ais_rx_status",
    "b200ais_rx_tag_overflows", "b200ais_selftest_div",
    "b200ais_rx_replay_file", "b200ais_rx_serve_udp",
]
FRAME_MAX = 248
FRAME_DTYPE = np.dtype([("end_bit", "<u8"), ("len", "<i4"), ("channel", "<i4"),
                        ("data", "u1", (FRAME_MAX,))])
E_FRAME_OVERFLOW = -8
STAGE_NAMES = ["sqfft_freqest", "nco_phase", "mix_agc", "corr", "detect", "msk", "tail"]


class DemodConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_float), ("data_rate", C.c_int), ("fftlen", C.c_int),
                ("agc_nsamples", C.c_int), ("agc_reference", C.c_float), ("sps", C.c_float),
                ("mark_delay", C.c_uint), ("threshold", C.c_float), ("gain", C.c_float),
                ("limit", C.c_float), ("osps", C.c_int), ("corr_chunk", C.c_int),
                ("stages", C.c_int)]


class RxConfig(C.Structure):
    _fields_ = [("rate", C.c_double), ("nfreqs", C.c_int), ("freqs", C.c_double * 16),
                ("designators", (C.c_char * 8) * 16), ("sources", C.c_int),
                ("max_input_items", C.c_int), ("max_frames", C.c_int), ("bits_per_sec", C.c_float),
                ("clockrec_gain", C.c_float), ("omega_relative_limit", C.c_float),
                ("fftlen", C.c_int), ("lpf_cutoff", C.c_double), ("lpf_transition", C.c_double),
                ("hdlc_length_min", C.c_int), ("hdlc_length_max", C.c_int)]


RX_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int)


class B200AisError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("b200ais error %d: %s" % (code, text))
        self.code = code


def build(force=False, verbose=False):
    """Compile libb200ais.so in-tree with nvcc for sm_100a (gr-ais_b200/csrc/Makefile)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run(["make", "-C", CSRC, "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    if verbose or r.returncode:
        print(r.stdout)
    if r.returncode:
        raise RuntimeError("building libb200ais.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load libb200ais.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libb200ais.so is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (needs nvcc); the demod path has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.b200ais_last_error.restype = C.c_char_p
    L.b200ais_launch_count.restype = C.c_uint64
    L.b200ais_corr_est_threshold.restype = C.c_float
    L.b200ais_corr_est_mark_delay.restype = C.c_uint
    for f in ("b200ais_msk_get_gain", "b200ais_msk_get_limit", "b200ais_msk_get_sps"):
        getattr(L, f).restype = C.c_float
    vp, i, u64, f32, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_float, C.c_size_t
    L.b200ais_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.b200ais_host_free.argtypes = [vp]
    L.b200ais_corr_est_create.argtypes = [C.POINTER(vp), vp, i, f32, C.c_uint, f32, i]
    L.b200ais_corr_est_destroy.argtypes = [vp]
    L.b200ais_corr_est_set_symbols.argtypes = [vp, vp, i]
    L.b200ais_corr_est_symbols.argtypes = [vp, vp, i, C.POINTER(i)]
    for f in ("b200ais_corr_est_output_multiple", "b200ais_corr_est_history",
              "b200ais_corr_est_mark_delay", "b200ais_corr_est_threshold"):
        getattr(L, f).argtypes = [vp]
    L.b200ais_corr_est_work.argtypes = [vp, i, vp, sz, u64, vp, vp, sz, vp, i, vp]
    L.b200ais_corr_est_work_dev.argtypes = [vp, i, vp, sz, u64, vp, vp, sz, vp, i, vp, vp]
    L.b200ais_msk_create.argtypes = [C.POINTER(vp), f32, f32, f32, i, i]
    L.b200ais_msk_destroy.argtypes = [vp]
    L.b200ais_msk_set_gain.argtypes = [vp, f32]
    L.b200ais_msk_set_limit.argtypes = [vp, f32]
    L.b200ais_msk_set_sps.argtypes = [vp, f32]
    for f in ("b200ais_msk_get_gain", "b200ais_msk_get_limit", "b200ais_msk_get_sps",
              "b200ais_msk_reset"):
        getattr(L, f).argtypes = [vp]
    L.b200ais_msk_forecast.argtypes = [vp, i]
    L.b200ais_msk_general_work.argtypes = [vp, i, i, vp, sz, u64, vp, i, vp, vp, vp, vp, sz, vp, vp]
    L.b200ais_msk_general_work_dev.argtypes = [vp, i, i, vp, sz, u64, vp, i, vp, vp, vp, vp, sz, vp,
                                               vp, vp]
    L.b200ais_freqest_create.argtypes = [C.POINTER(vp), f32, i, i, i]
    L.b200ais_freqest_destroy.argtypes = [vp]
    L.b200ais_freqest_work.argtypes = [vp, i, vp, vp]
    L.b200ais_freqest_work_dev.argtypes = [vp, i, vp, vp, vp]
    L.b200ais_invert_work.argtypes = [vp, vp, sz]
    L.b200ais_invert_work_dev.argtypes = [vp, vp, sz, vp]
    L.b200ais_demod_default_config.argtypes = [C.POINTER(DemodConfig)]
    L.b200ais_demod_create.argtypes = [C.POINTER(vp), C.POINTER(DemodConfig), vp, i, i, i, i]
    L.b200ais_demod_destroy.argtypes = [vp]
    L.b200ais_demod_max_bits.argtypes = [vp, i]
    L.b200ais_demod_work.argtypes = [vp, vp, i, vp, i, vp, vp, vp]
    L.b200ais_demod_work_sc16.argtypes = [vp, vp, C.c_float, i, vp, i, vp, vp, vp]
    L.b200ais_demod_work_dev.argtypes = [vp, vp, i, vp, i, vp, vp, vp, vp]
    L.b200ais_demod_status.argtypes = [vp]
    L.b200ais_demod_enable_taps.argtypes = [vp, i]
    L.b200ais_demod_tap.argtypes = [vp, i, C.POINTER(vp), C.POINTER(sz)]
    L.b200ais_demod_read_tap.argtypes = [vp, i, vp, sz]
    L.b200ais_demod_profile.argtypes = [vp, i]
    L.b200ais_demod_set_overlap.argtypes = [vp, i]
    L.b200ais_demod_stage_ms.argtypes = [vp, vp, vp]
    L.b200ais_demod_enqueue_dev.argtypes = [vp, vp, i, vp, i, vp, vp, vp, vp]
    L.b200ais_demod_join.argtypes = [vp, vp]
    L.b200ais_demod_set_symbols.argtypes = [vp, vp, i, vp]
    L.b200ais_demod_stream_reset.argtypes = [vp, vp]
    L.b200ais_demod_stream_max_bits.argtypes = [vp, i]
    L.b200ais_demod_stream_work.argtypes = [vp, vp, i, vp, i, vp, vp, vp]
    L.b200ais_demod_stream_work_dev.argtypes = [vp, vp, i, vp, i, vp, vp, vp, vp]
    L.b200ais_demod_stream_pending.argtypes = [vp, vp, vp, vp]
    f64 = C.c_double
    L.b200ais_firdes_low_pass.argtypes = [f64, f64, f64, f64, vp, i, C.POINTER(i)]
    L.b200ais_xlat_create.argtypes = [C.POINTER(vp), i, vp, i, vp, i, f64, i]
    for f in ("b200ais_xlat_destroy", "b200ais_xlat_history", "b200ais_xlat_decimation",
              "b200ais_xlat_reset"):
        getattr(L, f).argtypes = [vp]
    L.b200ais_xlat_set_center_freq.argtypes = [vp, i, f64]
    L.b200ais_xlat_set_taps.argtypes = [vp, vp, i]
    L.b200ais_xlat_work.argtypes = [vp, i, vp, sz, vp, sz]
    L.b200ais_xlat_work_dev.argtypes = [vp, i, vp, sz, vp, sz, vp]
    L.b200ais_hdlc_create.argtypes = [C.POINTER(vp), i, i, i]
    for f in ("b200ais_hdlc_destroy", "b200ais_hdlc_reset", "b200ais_hdlc_status"):
        getattr(L, f).argtypes = [vp]
    L.b200ais_hdlc_work.argtypes = [vp, vp, sz, vp, i, vp, i, vp]
    L.b200ais_hdlc_work_dev.argtypes = [vp, vp, sz, vp, i, vp, i, vp, vp]
    L.b200ais_nmea_slot_bytes.argtypes = [i, C.c_char_p]
    L.b200ais_nmea_format.argtypes = [vp, vp, i, i, vp, vp, i, vp]
    L.b200ais_nmea_format_dev.argtypes = [vp, vp, i, i, vp, vp, i, vp, vp]
    L.b200ais_demod_stream_stage.argtypes = [vp, i, i, C.POINTER(vp), C.POINTER(sz), vp]
    L.b200ais_demod_stream_work_staged.argtypes = [vp, i, vp, i, vp, vp, vp, vp]
    L.b200ais_rx_default_config.argtypes = [C.POINTER(RxConfig)]
    L.b200ais_rx_create.argtypes = [C.POINTER(vp), C.POINTER(RxConfig), vp, i]
    for f in ("b200ais_rx_destroy", "b200ais_rx_reset", "b200ais_rx_decimation",
              "b200ais_rx_channels", "b200ais_rx_samples_per_symbol", "b200ais_rx_sentence_slot",
              "b200ais_rx_status"):
        getattr(L, f).argtypes = [vp]
    L.b200ais_rx_samples_per_symbol.restype = C.c_float
    L.b200ais_rx_tag_overflows.argtypes = [vp]
    L.b200ais_rx_tag_overflows.restype = u64
    L.b200ais_selftest_div.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.POINTER(C.c_ulonglong)]
    L.b200ais_rx_work.argtypes = [vp, vp, sz, i, vp, vp, i, vp, i, C.POINTER(i)]
    L.b200ais_rx_work_dev.argtypes = [vp, vp, sz, i, vp, vp, i, vp, i, vp, vp]
    L.b200ais_rx_replay_file.argtypes = [vp, C.c_char_p, i, i, RX_SINK, vp, C.POINTER(u64)]
    L.b200ais_rx_serve_udp.argtypes = [vp, C.c_char_p, i, i, i, u64, i, RX_SINK, vp, C.POINTER(u64)]
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise B200AisError(rc, lib().b200ais_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Raw address of a numpy array, torch tensor (host or CUDA), int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def device_count():
    n = C.c_int(0)
    check(lib().b200ais_device_count(C.byref(n)))
    return n.value


def set_device(d):
    check(lib().b200ais_set_device(int(d)))


def launch_count():
    return int(lib().b200ais_launch_count())


def default_config(**over):
    cfg = DemodConfig()
    check(lib().b200ais_demod_default_config(C.byref(cfg)))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


class PinnedArray:
    """numpy view of page-locked host memory from b200ais_host_alloc."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = C.c_void_p()
        check(lib().b200ais_host_alloc(C.byref(self._p), max(nbytes, 1)))
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p and self._p.value:
            self.array = None
            check(lib().b200ais_host_free(self._p))
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
