"""ctypes binding of the CPU oracle (oracle/libais_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libais_oracle.so")

TAG_CORR_START, TAG_PHASE_EST, TAG_TIME_EST, TAG_CORR_EST = 0, 1, 2, 3
STAGE_FREQSYNC, STAGE_AGC = 1, 2

TAG_DTYPE = np.dtype([("offset", "<u8"), ("key", "<i4"), ("port", "<i4"), ("value", "<f8")])


class Tag(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("key", C.c_int32), ("port", C.c_int32),
                ("value", C.c_double)]


class FftFilt(C.Structure):
    _fields_ = [("ntaps", C.c_int), ("fftsize", C.c_int), ("nsamples", C.c_int),
                ("H", C.POINTER(C.c_float)), ("tail", C.POINTER(C.c_float)),
                ("tw", C.POINTER(C.c_float))]


class CorrEst(C.Structure):
    _fields_ = [("taps", C.POINTER(C.c_float)), ("L", C.c_int), ("sps", C.c_float),
                ("mark_delay", C.c_uint), ("thresh", C.c_float), ("f", FftFilt)]


class Msk(C.Structure):
    _fields_ = [("sps", C.c_float), ("gain", C.c_float), ("gain_omega", C.c_float),
                ("limit", C.c_float), ("mu", C.c_float), ("omega", C.c_float),
                ("dly1_re", C.c_float), ("dly1_im", C.c_float), ("dly2_re", C.c_float),
                ("dly2_im", C.c_float), ("diff1_re", C.c_float), ("diff1_im", C.c_float),
                ("div", C.c_int), ("osps", C.c_int), ("prev_re", C.c_float),
                ("prev_im", C.c_float)]


class FreqEst(C.Structure):
    _fields_ = [("offset", C.c_int), ("binsize", C.c_float), ("fftlen", C.c_int)]


class ChainCfg(C.Structure):
    _fields_ = [("sample_rate", C.c_float), ("data_rate", C.c_int), ("fftlen", C.c_int),
                ("agc_nsamples", C.c_int), ("agc_reference", C.c_float), ("sps", C.c_float),
                ("mark_delay", C.c_uint), ("threshold", C.c_float), ("gain", C.c_float),
                ("limit", C.c_float), ("osps", C.c_int), ("corr_chunk", C.c_int),
                ("stages", C.c_int)]


class ChainOut(C.Structure):
    _fields_ = [("bits", C.c_void_p), ("max_bits", C.c_int), ("nbits", C.c_int),
                ("tags", C.c_void_p), ("max_tags", C.c_int), ("ntags", C.c_int),
                ("fhat", C.c_void_p), ("mixed", C.c_void_p), ("agc", C.c_void_p),
                ("corr", C.c_void_p), ("mag", C.c_void_p), ("sym", C.c_void_p),
                ("err", C.c_void_p), ("mu", C.c_void_p), ("soft", C.c_void_p),
                ("n1", C.c_int), ("n2", C.c_int), ("consumed", C.c_int)]


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc only; seconds)."""
    src = [os.path.join(_HERE, f) for f in ("ais_oracle.c", "ais_oracle_rx.c", "ais_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ao_fast_atan2f.restype = C.c_float
        _lib.ao_fast_atan2f.argtypes = [C.c_float, C.c_float]
        _lib.ao_hypotf.restype = C.c_float
        _lib.ao_hypotf.argtypes = [C.c_float, C.c_float]
        _lib.ao_branchless_clip.restype = C.c_float
        _lib.ao_branchless_clip.argtypes = [C.c_float, C.c_float]
        _lib.ao_float_to_fixed.restype = C.c_int32
        _lib.ao_float_to_fixed.argtypes = [C.c_float]
        _lib.ao_agc_envelope.restype = C.c_float
        _lib.ao_agc_envelope.argtypes = [C.c_float, C.c_float]
        for name in ("ao_mmse_taps", "ao_atan_table", "ao_sine_table"):
            getattr(_lib, name).restype = C.POINTER(C.c_float)
        _lib.ao_default_corr_chunk.restype = C.c_int
        _lib.ao_default_corr_chunk.argtypes = [C.c_int]
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def _c64(a):
    a = np.ascontiguousarray(a, dtype=np.complex64)
    return a


# ---------------------------------------------------------------- tables

def mmse_taps():
    return np.ctypeslib.as_array(lib().ao_mmse_taps(), shape=(129, 8)).copy()


def atan_table():
    return np.ctypeslib.as_array(lib().ao_atan_table(), shape=(257,)).copy()


def sine_table():
    return np.ctypeslib.as_array(lib().ao_sine_table(), shape=(1024, 2)).copy()


# --------------------------------------------------------------- scalars

def fast_atan2f(y, x):
    return float(lib().ao_fast_atan2f(float(np.float32(y)), float(np.float32(x))))


def hypotf(re, im):
    return float(lib().ao_hypotf(float(np.float32(re)), float(np.float32(im))))


def branchless_clip(x, c):
    return float(lib().ao_branchless_clip(float(np.float32(x)), float(np.float32(c))))


def float_to_fixed(x):
    return int(lib().ao_float_to_fixed(float(np.float32(x))))


def fxpt_sincos(angle):
    s, c = C.c_float(), C.c_float()
    lib().ao_fxpt_sincos(C.c_int32(angle), C.byref(s), C.byref(c))
    return s.value, c.value


def agc_envelope(re, im):
    return float(lib().ao_agc_envelope(float(np.float32(re)), float(np.float32(im))))


# -------------------------------------------------------------- template

def gmsk_template_packed(data, sps=5, bt=0.4):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(len(data) * 8 * sps, dtype=np.complex64)
    n = lib().ao_gmsk_template_packed(_fp(data), len(data), int(sps), C.c_float(bt), _fp(out))
    assert n == len(out)
    return out


def gmsk_template_bits(bits, sps=5, bt=0.4):
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(len(bits) * sps, dtype=np.complex64)
    n = lib().ao_gmsk_template_bits(_fp(bits), len(bits), int(sps), C.c_float(bt), _fp(out))
    assert n == len(out)
    return out


def firdes_gaussian(gain, spb, bt, ntaps):
    out = np.zeros(ntaps, dtype=np.float32)
    lib().ao_firdes_gaussian(C.c_double(gain), C.c_double(spb), C.c_double(bt), int(ntaps), _fp(out))
    return out


# ------------------------------------------------------------ freq sync

def square(x):
    x = _c64(x)
    out = np.empty_like(x)
    lib().ao_square(_fp(x), _fp(out), len(x))
    return out


def fft_forward(x):
    x = _c64(x)
    out = np.empty_like(x)
    rc = lib().ao_fft_forward(_fp(x), _fp(out), len(x))
    if rc:
        raise ValueError("fft length must be a power of two >= 2")
    return out


def fft_dif(x):
    """correlator forward transform: natural-order in, bit-reversed out"""
    x = _c64(x).copy()
    if lib().ao_fft_dif_inplace(_fp(x), len(x)):
        raise ValueError("fft length must be a power of two >= 2")
    return x


def ifft_dit(x):
    """correlator inverse transform (unnormalised): bit-reversed in, natural-order out"""
    x = _c64(x).copy()
    if lib().ao_ifft_dit_inplace(_fp(x), len(x)):
        raise ValueError("fft length must be a power of two >= 2")
    return x


def fft_shift(x):
    x = _c64(x)
    out = np.empty_like(x)
    lib().ao_fft_shift(_fp(x), _fp(out), len(x))
    return out


def freqest_work(spec, sample_rate=48000.0, data_rate=9600, fftlen=1024):
    """spec: [nvec, fftlen] complex64 -> (hz[nvec] float32, maxpos[nvec] int32)"""
    spec = _c64(spec).reshape(-1, fftlen)
    fe = FreqEst()
    lib().ao_freqest_init(C.byref(fe), C.c_float(sample_rate), int(data_rate), int(fftlen))
    out = np.zeros(len(spec), dtype=np.float32)
    mp = np.zeros(len(spec), dtype=np.int32)
    lib().ao_freqest_work(C.byref(fe), _fp(spec), len(spec), _fp(out), _fp(mp))
    return out, mp


def nco_mix(x, freq, rep, sensitivity, phase=0.0):
    x = _c64(x)
    freq = np.ascontiguousarray(freq, dtype=np.float32)
    out = np.empty_like(x)
    ph = C.c_float(phase)
    lib().ao_nco_mix(C.byref(ph), C.c_float(sensitivity), _fp(freq), int(rep), _fp(x), len(x),
                     _fp(out))
    return out, ph.value


def agc_work(x, nsamples=512, reference=2.0, history=None):
    """x: stream; history (nsamples-1 items) defaults to zeros.  Returns len(x) items."""
    x = _c64(x)
    if history is None:
        history = np.zeros(nsamples - 1, dtype=np.complex64)
    buf = np.concatenate([_c64(history), x])
    out = np.empty(len(x), dtype=np.complex64)
    lib().ao_agc_work(_fp(buf), len(x), int(nsamples), C.c_float(reference), _fp(out))
    return out


# ------------------------------------------------------------- corr_est

class CorrEstBlock:
    """Mirror of gr::ais::corr_est_cc driven one work() call at a time."""

    def __init__(self, symbols, sps, mark_delay, threshold=0.9):
        symbols = _c64(symbols)
        self._c = CorrEst()
        rc = lib().ao_corr_est_init(C.byref(self._c), _fp(symbols), len(symbols), C.c_float(sps),
                                    C.c_uint(mark_delay), C.c_float(threshold))
        assert rc == 0

    def __del__(self):
        try:
            lib().ao_corr_est_free(C.byref(self._c))
        except Exception:
            pass

    @property
    def L(self):
        return self._c.L

    @property
    def thresh(self):
        return self._c.thresh

    @property
    def nsamples(self):
        return self._c.f.nsamples

    @property
    def mark_delay(self):
        return self._c.mark_delay

    def symbols(self):
        return np.ctypeslib.as_array(self._c.taps, shape=(2 * self._c.L,)).copy().view(np.complex64)

    def set_symbols(self, symbols):
        symbols = _c64(symbols)
        lib().ao_corr_est_set_symbols(C.byref(self._c), _fp(symbols), len(symbols))

    def work(self, n, inbuf, nitems_written=0, two_ports=False, max_tags=4096):
        """inbuf: n + L items (history first).  Returns out0, corr, mag, tags."""
        inbuf = _c64(inbuf)
        assert len(inbuf) >= n + self._c.L
        out0 = np.empty(n, dtype=np.complex64)
        corr = np.empty(n, dtype=np.complex64)
        mag = np.empty(n, dtype=np.float32)
        tags = np.zeros(max_tags, dtype=TAG_DTYPE)
        nt = C.c_int(0)
        rc = lib().ao_corr_est_work(C.byref(self._c), int(n), _fp(inbuf), C.c_uint64(nitems_written),
                                    _fp(out0), _fp(corr), _fp(mag), int(bool(two_ports)), _fp(tags),
                                    int(max_tags), C.byref(nt))
        if rc < 0:
            raise ValueError("noutput_items must be a multiple of the output multiple (%d)" % self.nsamples)
        if nt.value > max_tags:
            raise RuntimeError("tag buffer overflow")
        return out0, corr, mag, tags[:nt.value].copy()

    @property
    def fftsize(self):
        return self._c.f.fftsize

    def tail(self):
        n = max(self._c.L - 1, 0)
        return np.ctypeslib.as_array(self._c.f.tail, shape=(2 * max(n, 1),)).copy().view(np.complex64)[:n]

    def direct_f64(self, n, inbuf):
        """float64 direct-form correlation of the same call (truth for tests; ignores the tail)"""
        inbuf = _c64(inbuf)
        out = np.zeros(n, dtype=np.complex128)
        lib().ao_corr_direct_f64(C.byref(self._c), int(n), _fp(inbuf), _fp(out))
        return out


# ------------------------------------------------------------------ msk

class MskBlock:
    """Mirror of gr::ais::msk_timing_recovery_cc driven one general_work() at a time."""

    def __init__(self, sps, gain, limit, osps=1):
        self._m = Msk()
        rc = lib().ao_msk_init(C.byref(self._m), C.c_float(sps), C.c_float(gain), C.c_float(limit),
                               int(osps))
        if rc == -1:
            raise IndexError("Gain must be positive")  # std::out_of_range
        if rc == -2:
            raise IndexError("osps must be 1 or 2")

    @property
    def state(self):
        return self._m

    def forecast(self, noutput_items):
        return int(lib().ao_msk_forecast(C.byref(self._m), int(noutput_items)))

    def general_work(self, noutput_items, inbuf, tags=None, nitems_read=0):
        inbuf = _c64(inbuf)
        if tags is None:
            tags = np.zeros(0, dtype=TAG_DTYPE)
        tags = np.ascontiguousarray(tags, dtype=TAG_DTYPE)
        out = np.empty(max(noutput_items, 1), dtype=np.complex64)
        err = np.empty(max(noutput_items, 1), dtype=np.float32)
        mu = np.empty(max(noutput_items, 1), dtype=np.float32)
        consumed = C.c_int(0)
        k = lib().ao_msk_general_work(C.byref(self._m), int(noutput_items), len(inbuf), _fp(inbuf),
                                      C.c_uint64(nitems_read), _fp(tags), len(tags), _fp(out),
                                      _fp(err), _fp(mu), C.byref(consumed))
        if k < 0:
            raise RuntimeError("mmse interpolator index out of range")
        return out[:k].copy(), err[:k].copy(), mu[:k].copy(), consumed.value


# ----------------------------------------------------------------- tail

def quad_demod(x, gain=float(np.float32(np.pi / 2)), prev=0j):
    x = _c64(x)
    out = np.empty(len(x), dtype=np.float32)
    p = np.array([prev], dtype=np.complex64)
    lib().ao_quad_demod(_fp(p), _fp(x), len(x), C.c_float(gain), _fp(out))
    return out


def binary_slicer(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(len(x), dtype=np.uint8)
    lib().ao_binary_slicer(_fp(x), len(x), _fp(out))
    return out


def diff_decoder(b, modulus=2, prev=0):
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(len(b), dtype=np.uint8)
    p = C.c_uint8(prev)
    lib().ao_diff_decoder(C.byref(p), _fp(b), len(b), C.c_uint(modulus), _fp(out))
    return out


def invert(b):
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(len(b), dtype=np.uint8)
    lib().ao_invert(_fp(b), len(b), _fp(out))
    return out


# ---------------------------------------------------------------- chain

def chain_cfg(sample_rate=48000.0, data_rate=9600, fftlen=1024, agc_nsamples=512,
              agc_reference=2.0, sps=5.0, mark_delay=1, threshold=0.9, gain=0.04, limit=0.01,
              osps=1, corr_chunk=0, stages=STAGE_FREQSYNC | STAGE_AGC):
    return ChainCfg(sample_rate, data_rate, fftlen, agc_nsamples, agc_reference, sps, mark_delay,
                    threshold, gain, limit, osps, corr_chunk, stages)


def default_corr_chunk(L):
    return int(lib().ao_default_corr_chunk(int(L)))


def max_bits_for(n, sps=5.0, osps=1):
    return int(n / sps * osps * 1.05) + 64


def demod_chain(x, symbols, cfg=None, debug=False, max_tags=4096, blocks=None):
    """One record from fresh state.  Returns dict(bits, tags, [debug taps]).
    blocks: ao_blocks pointer (oracle.ref.blocks() = the reference's own classes); None = the
    restated blocks."""
    cfg = cfg or chain_cfg()
    x = _c64(x)
    symbols = _c64(symbols)
    n = len(x)
    mb = max_bits_for(n, cfg.sps, cfg.osps)
    bits = np.zeros(mb, dtype=np.uint8)
    tags = np.zeros(max_tags, dtype=TAG_DTYPE)
    o = ChainOut()
    o.bits, o.max_bits, o.tags, o.max_tags = _fp(bits).value, mb, _fp(tags).value, max_tags
    dbg = {}
    if debug:
        dbg = dict(fhat=np.zeros(n // cfg.fftlen + 1, np.float32), mixed=np.zeros(n, np.complex64),
                   agc=np.zeros(n, np.complex64), corr=np.zeros(n, np.complex64),
                   mag=np.zeros(n, np.float32), sym=np.zeros(mb, np.complex64),
                   err=np.zeros(mb, np.float32), mu=np.zeros(mb, np.float32),
                   soft=np.zeros(mb, np.float32))
        for k, v in dbg.items():
            setattr(o, k, _fp(v).value)
    lib().ao_demod_chain_with.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int, C.c_void_p]
    rc = lib().ao_demod_chain_with(blocks, C.addressof(cfg), _fp(symbols), len(symbols), _fp(x), n,
                                   C.addressof(o))
    if rc:
        raise RuntimeError("ao_demod_chain failed: %d" % rc)
    res = dict(bits=bits[:o.nbits].copy(), tags=tags[:o.ntags].copy(), n1=o.n1, n2=o.n2,
               consumed=o.consumed)
    if debug:
        nv = o.n1 // cfg.fftlen if (cfg.stages & STAGE_FREQSYNC) else 0
        res.update(fhat=dbg["fhat"][:nv], mixed=dbg["mixed"][:o.n1], agc=dbg["agc"][:o.n1],
                   corr=dbg["corr"][:o.n2], mag=dbg["mag"][:o.n2], sym=dbg["sym"][:o.nbits],
                   err=dbg["err"][:o.nbits], mu=dbg["mu"][:o.nbits], soft=dbg["soft"][:o.nbits])
    return res


class DemodStream:
    """The chain as a stream (ao_stream): state carried from work() call to work() call."""

    def __init__(self, symbols, cfg=None, blocks=None):
        self.cfg = cfg or chain_cfg()
        symbols = _c64(symbols)
        L = lib()
        L.ao_stream_new_with.restype = C.c_void_p
        L.ao_stream_new_with.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ao_stream_delete.argtypes = [C.c_void_p]
        L.ao_stream_work.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]
        self._s = L.ao_stream_new_with(blocks, C.addressof(self.cfg), _fp(symbols), len(symbols))
        if not self._s:
            raise RuntimeError("ao_stream_new failed")

    def __del__(self):
        try:
            if self._s:
                lib().ao_stream_delete(self._s)
                self._s = None
        except Exception:
            pass

    def set_symbols(self, symbols):
        symbols = _c64(symbols)
        lib().ao_stream_set_symbols.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        if lib().ao_stream_set_symbols(self._s, _fp(symbols), len(symbols)):
            raise ValueError("set_symbols: the stream was built for another template length")

    def work(self, x, max_tags=4096):
        """Returns (bits, tags) produced by this call."""
        x = _c64(x)
        mb = max_bits_for(len(x) + 2 * self.cfg.fftlen + 4096, self.cfg.sps, self.cfg.osps)
        bits = np.zeros(mb, dtype=np.uint8)
        tags = np.zeros(max_tags, dtype=TAG_DTYPE)
        nb, nt = C.c_int(0), C.c_int(0)
        rc = lib().ao_stream_work(self._s, _fp(x), len(x), _fp(bits), mb, C.byref(nb), _fp(tags), max_tags,
                                  C.byref(nt))
        if rc:
            raise RuntimeError("ao_stream_work failed: %d" % rc)
        return bits[:nb.value].copy(), tags[:nt.value].copy()


def demod_chain_batch(x, symbols, cfg=None, max_tags=256, nthreads=0, blocks=None):
    """x: [C, n] complex64.  Returns bits [C, max_bits], nbits [C], tags [C, max_tags], ntags [C]."""
    cfg = cfg or chain_cfg()
    x = np.ascontiguousarray(x, dtype=np.complex64)
    symbols = _c64(symbols)
    Cn, n = x.shape
    mb = max_bits_for(n, cfg.sps, cfg.osps)
    bits = np.zeros((Cn, mb), dtype=np.uint8)
    nbits = np.zeros(Cn, dtype=np.int32)
    tags = np.zeros((Cn, max_tags), dtype=TAG_DTYPE)
    ntags = np.zeros(Cn, dtype=np.int32)
    lib().ao_demod_chain_batch_with.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    rc = lib().ao_demod_chain_batch_with(blocks, C.addressof(cfg), _fp(symbols), len(symbols), _fp(x),
                                         Cn, n, _fp(bits), mb, _fp(nbits), _fp(tags), max_tags,
                                         _fp(ntags), int(nthreads))
    if rc:
        raise RuntimeError("ao_demod_chain_batch failed: %d" % rc)
    return bits, nbits, tags, ntags


# ------------------------------------------- rx rows (ais_oracle_rx.c)

FRAME_MAX = 248
FRAME_DTYPE = np.dtype([("end_bit", "<u8"), ("len", "<i4"), ("channel", "<i4"),
                        ("data", "u1", (FRAME_MAX,))])


class Hdlc(C.Structure):
    _fields_ = [("length_min", C.c_int), ("length_max", C.c_int), ("ones", C.c_int),
                ("bitctr", C.c_int), ("bytectr", C.c_int), ("pad", C.c_int),
                ("nitems_read", C.c_uint64), ("pktbuf", C.c_uint8 * (FRAME_MAX + 8))]


class Xlat(C.Structure):
    _fields_ = [("decim", C.c_int), ("ntaps", C.c_int), ("ctaps", C.POINTER(C.c_float)),
                ("incr_re", C.c_float), ("incr_im", C.c_float), ("phase_re", C.c_float),
                ("phase_im", C.c_float), ("counter", C.c_uint)]


def crc_ccitt(data: bytes) -> int:
    L = lib()
    L.ao_crc_ccitt.restype = C.c_uint
    L.ao_crc_ccitt.argtypes = [C.c_char_p, C.c_size_t]
    return int(L.ao_crc_ccitt(bytes(data), len(data)))


class HdlcDeframer:
    """digital.hdlc_deframer_bp(length_min, length_max) [G], streaming."""

    def __init__(self, length_min=11, length_max=64):
        self._h = Hdlc()
        lib().ao_hdlc_init(C.byref(self._h), int(length_min), int(length_max))

    def work(self, bits, max_frames=256):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        frames = np.zeros(max_frames, dtype=FRAME_DTYPE)
        dropped = C.c_int(0)
        L = lib()
        L.ao_hdlc_work.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p]
        n = L.ao_hdlc_work(C.byref(self._h), _fp(bits), len(bits), _fp(frames), max_frames,
                           C.byref(dropped))
        if dropped.value:
            raise OverflowError("more than max_frames frames")
        return frames[:n]


def frames_payloads(frames):
    return [bytes(f["data"][:f["len"]]) for f in frames]


def pdu_to_nmea(data: bytes, designator="A") -> str:
    L = lib()
    L.ao_pdu_to_nmea.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(4096)
    n = L.ao_pdu_to_nmea(designator.encode(), bytes(data), len(data), buf, 4096)
    if n < 0:
        raise ValueError("pdu_to_nmea: bad length")
    return buf.raw[:n].decode("latin-1")


def firdes_low_pass(gain, fs, cutoff, tw):
    L = lib()
    L.ao_firdes_low_pass.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                     C.c_int]
    n = L.ao_firdes_low_pass(gain, fs, cutoff, tw, None, 0)
    taps = np.zeros(n, dtype=np.float32)
    assert L.ao_firdes_low_pass(gain, fs, cutoff, tw, _fp(taps), n) == n
    return taps


class FreqXlatingFir:
    """filter.freq_xlating_fir_filter_ccf(decimation, taps, center_freq, sampling_freq) [G]."""

    def __init__(self, decimation, taps, center_freq, sampling_freq):
        self.taps = np.ascontiguousarray(taps, dtype=np.float32)
        self.decim = int(decimation)
        self.center_freq, self.sampling_freq = float(center_freq), float(sampling_freq)
        self._x = Xlat()
        L = lib()
        L.ao_xlat_init.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double]
        if L.ao_xlat_init(C.byref(self._x), self.decim, _fp(self.taps), len(self.taps),
                          self.center_freq, self.sampling_freq):
            raise ValueError("xlat_init")
        self.nout = 0

    def __del__(self):
        try:
            lib().ao_xlat_free(C.byref(self._x))
        except Exception:
            pass

    @property
    def ctaps(self):
        return np.ctypeslib.as_array(self._x.ctaps, shape=(2 * len(self.taps),)).copy().view(np.complex64)

    def work(self, inbuf, fir=False):
        """inbuf: ntaps-1 history items followed by noutput*decim new items."""
        inbuf = _c64(inbuf)
        n = (len(inbuf) - (len(self.taps) - 1)) // self.decim
        out = np.zeros(n, dtype=np.complex64)
        fo = np.zeros(n, dtype=np.complex64) if fir else None
        L = lib()
        L.ao_xlat_work.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ao_xlat_work(C.byref(self._x), n, _fp(inbuf), _fp(out), _fp(fo) if fir else None)
        self.nout += n
        return (out, fo) if fir else out

    def f64(self, inbuf, first_output=0):
        inbuf = _c64(inbuf)
        n = (len(inbuf) - (len(self.taps) - 1)) // self.decim
        out = np.zeros(2 * n, dtype=np.float64)
        L = lib()
        L.ao_xlat_f64.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_uint64,
                                  C.c_int, C.c_void_p, C.c_void_p]
        L.ao_xlat_f64(self.decim, _fp(self.taps), len(self.taps), self.center_freq,
                      self.sampling_freq, int(first_output), n, _fp(inbuf), _fp(out))
        return out.view(np.complex128)
