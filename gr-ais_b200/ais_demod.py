"""Host-side mirrors of the reference's two Python hier-blocks.

  ais_demod(options)                       reference python/ais_demod.py:21-56
  square_and_fft_sync_cc(rate, bps, fftlen) reference python/gmsk_sync.py:14-37

The reference wires stock GNU Radio blocks around its own; here the whole wiring is
one fused CUDA chain (b200ais_demod_* in include/b200ais.h) and these classes only
carry the same construction arguments: the options dict of python/radio.py:56-62.
"""
import ctypes as C
import math
import os

import numpy as np

from . import binding as B

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "tables")
_f32 = np.float32


def _read_table(name, shape):
    vals = []
    with open(os.path.join(_TABLES, name)) as fh:
        for line in fh:
            line = line.split("/*")[0]
            for tok in line.replace("{", " ").replace("}", " ").replace(",", " ").split():
                vals.append(float(tok))
    return np.array(vals, dtype=np.float64).astype(np.float32).reshape(shape)


_SINE = None


def _fxpt_sincos(angle):
    """gr::fxpt::sincos on an int32 angle -> (sin, cos) float32."""
    global _SINE
    if _SINE is None:
        _SINE = _read_table("sine_table.inc", (1024, 2))
    ux = angle & 0xFFFFFFFF
    e = _SINE[ux >> 22]
    s = _f32(e[0] * _f32(ux >> 1)) + e[1]
    ux = (angle + 0x40000000) & 0xFFFFFFFF
    e = _SINE[ux >> 22]
    c = _f32(e[0] * _f32(ux >> 1)) + e[1]
    return _f32(s), _f32(c)


def _float_to_fixed(x):
    """gr::fxpt::float_to_fixed (float32 arithmetic)."""
    pi = _f32(math.pi)
    two_pi = _f32(2.0) * pi
    d = int(math.floor(float(_f32(x) / two_pi) + 0.5))
    x = _f32(x) - _f32(_f32(d) * two_pi)
    v = _f32(_f32(x * _f32(2147483648.0)) / pi)
    if not (-2147483904.0 < float(v) < 2147483648.0):
        return -2 ** 31
    return int(v)  # truncation toward zero


def firdes_gaussian(gain, spb, bt, ntaps):
    """gr::filter::firdes::gaussian."""
    dt = 1.0 / spb
    s = 1.0 / (math.sqrt(math.log(2.0)) / (2 * math.pi * bt))
    t0 = -0.5 * ntaps
    taps = np.zeros(ntaps, dtype=np.float32)
    scale = 0.0
    for i in range(ntaps):
        t0 += 1
        ts = s * dt * t0
        taps[i] = math.exp(-0.5 * ts * ts)
        scale += float(taps[i])
    for i in range(ntaps):
        taps[i] = float(taps[i]) / scale * gain
    return taps


def gmsk_mod_bits(bits, samples_per_symbol=5, bt=0.4):
    """digital.gmsk_mod(sps, bt) run once over `bits` from zero filter/phase state: what
    digital.modulate_vector_bc(mod, data, [1]) returns (python/ais_demod.py:36-38)."""
    sps = int(samples_per_symbol)
    bits = np.asarray(bits, dtype=np.uint8)
    g = firdes_gaussian(1.0, sps, bt, 4 * sps)
    taps = np.convolve(g.astype(np.float64), np.ones(sps)).astype(np.float32)
    sens = _f32((math.pi / 2) / sps)
    pi = _f32(math.pi)
    phase = _f32(0.0)
    out = np.zeros(len(bits) * sps, dtype=np.complex64)
    for n in range(len(bits) * sps):
        acc = _f32(0.0)
        for k in range(n % sps, len(taps), sps):
            if n - k < 0:
                continue
            sym = (n - k) // sps
            if sym >= len(bits):
                continue
            v = _f32(1.0) if bits[sym] else _f32(-1.0)
            acc = _f32(acc + _f32(taps[k] * v))
        phase = _f32(phase + _f32(sens * acc))
        phase = _f32(_f32(np.fmod(_f32(phase + pi), _f32(_f32(2.0) * pi))) - pi)
        s, c = _fxpt_sincos(_float_to_fixed(phase))
        out[n] = complex(float(c), float(s))
    return out


def gmsk_mod_packed(data, samples_per_symbol=5, bt=0.4):
    """gmsk_mod consumes *packed* bytes, MSB first (reference include/ais/modulate_vector.h:52-56)."""
    bits = np.unpackbits(np.asarray(data, dtype=np.uint8), bitorder="big")
    return gmsk_mod_bits(bits, samples_per_symbol, bt)


REFERENCE_PREAMBLE = [1, 1, 0, 0] * 7  # python/ais_demod.py:36


def preamble_template(kind="north_star", sps=5, bt=0.4):
    """The corr_est_cc symbol template.

    "reference": python/ais_demod.py:36-38 literally -- the 28-entry list is handed to
                 gmsk_mod as packed bytes => 224 bits => 1120 taps at sps 5.
    "intended":  the same 28 entries taken as bits => 140 taps.
    "north_star": the 24-bit AIS training pattern [1,1,0,0]*6 as bits => 120 taps
                 (BASELINE.json north_star).
    """
    if kind == "reference":
        return gmsk_mod_packed(REFERENCE_PREAMBLE, sps, bt)
    if kind == "intended":
        return gmsk_mod_bits(REFERENCE_PREAMBLE, sps, bt)
    if kind == "north_star":
        return gmsk_mod_bits([1, 1, 0, 0] * 6, sps, bt)
    raise ValueError(kind)


def default_options():
    """The options dict ais_rx builds (python/radio.py:56-62) for a 48 ksps channel."""
    return {"samples_per_symbol": 5, "clockrec_gain": 0.04, "omega_relative_limit": 0.01,
            "bits_per_sec": 9600.0, "fftlen": 1024, "samp_rate": 48000.0}


class ais_demod:
    """complex baseband in, unpacked NRZI-decoded bits out (python/ais_demod.py:21-56).

    channels / max_samples size the device buffers; template selects the corr_est symbols
    (an array, or a preamble_template() kind).  stages lets a caller drop the freq-sync
    and/or AGC stages (BASELINE.json configs[1] runs corr_est + msk_timing only)."""

    def __init__(self, options=None, channels=1, max_samples=48000, template="north_star",
                 max_tags=256, stages=B.STAGE_FREQSYNC | B.STAGE_AGC, corr_chunk=0,
                 threshold=0.9, mark_delay=1, agc=(512, 2.0), osps=1):
        options = dict(default_options(), **(options or {}))
        self._samples_per_symbol = options["samples_per_symbol"]
        self._bits_per_sec = options["bits_per_sec"]
        self._samplerate = self._samples_per_symbol * self._bits_per_sec
        self._clockrec_gain = options["clockrec_gain"]
        self._omega_relative_limit = options["omega_relative_limit"]
        self.fftlen = options["fftlen"]
        if isinstance(template, str):
            template = preamble_template(template, int(self._samples_per_symbol))
        self.mod_vector = np.ascontiguousarray(template, dtype=np.complex64)
        self.channels = int(channels)
        self.max_samples = int(max_samples)
        self.max_tags = int(max_tags)
        self.cfg = B.default_config(sample_rate=float(self._samplerate),
                                    data_rate=int(self._bits_per_sec), fftlen=int(self.fftlen),
                                    sps=float(self._samples_per_symbol),
                                    gain=float(self._clockrec_gain),
                                    limit=float(self._omega_relative_limit),
                                    threshold=float(threshold), mark_delay=int(mark_delay),
                                    corr_chunk=int(corr_chunk), stages=int(stages),
                                    agc_nsamples=int(agc[0]), agc_reference=float(agc[1]), osps=int(osps))
        self._h = C.c_void_p()
        rc = B.lib().b200ais_demod_create(C.byref(self._h), C.byref(self.cfg), B.ptr(self.mod_vector),
                                          len(self.mod_vector), self.channels, self.max_samples,
                                          self.max_tags)
        if rc == B.E_RANGE:
            raise IndexError(B.lib().b200ais_last_error().decode())
        B.check(rc)

    def __del__(self):
        self.close()

    def close(self):
        try:
            if self._h:
                B.lib().b200ais_demod_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def max_bits(self, nsamples):
        return B.lib().b200ais_demod_max_bits(self._h, int(nsamples))

    def enable_taps(self, on=True):
        B.check(B.lib().b200ais_demod_enable_taps(self._h, 1 if on else 0))

    def work(self, iq, bits=None, nbits=None, tags=None, ntags=None):
        """iq: [channels, nsamples] complex64 host array (numpy, or pinned).  Returns
        (bits [channels, max_bits] uint8, nbits [channels], tags, ntags)."""
        iq = np.asarray(iq)
        if iq.dtype != np.complex64 or not iq.flags.c_contiguous:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim == 1:
            iq = iq.reshape(1, -1)
        if iq.shape[0] != self.channels:
            raise ValueError("expected %d channels" % self.channels)
        n = iq.shape[1]
        mb = self.max_bits(n) if bits is None else bits.shape[1]
        if bits is None:
            bits = np.zeros((self.channels, mb), dtype=np.uint8)
        if nbits is None:
            nbits = np.zeros(self.channels, dtype=np.int32)
        if tags is None:
            tags = np.zeros((self.channels, self.max_tags), dtype=B.TAG_DTYPE)
        if ntags is None:
            ntags = np.zeros(self.channels, dtype=np.int32)
        B.check(B.lib().b200ais_demod_work(self._h, B.ptr(iq), n, B.ptr(bits), mb, B.ptr(nbits),
                                           B.ptr(tags), B.ptr(ntags)))
        return bits, nbits, tags, ntags

    def work_sc16(self, iq16, scale, bits=None, nbits=None, tags=None, ntags=None):
        """work() with the IQ as interleaved int16: iq16 [channels, nsamples, 2] int16 host array;
        every component becomes float(v) * scale on the device (b200ais_demod_work_sc16)."""
        iq16 = np.asarray(iq16)
        if iq16.dtype != np.int16 or not iq16.flags.c_contiguous:
            iq16 = np.ascontiguousarray(iq16, dtype=np.int16)
        if iq16.ndim != 3 or iq16.shape[0] != self.channels or iq16.shape[2] != 2:
            raise ValueError("expected [%d, nsamples, 2] int16" % self.channels)
        n = iq16.shape[1]
        mb = self.max_bits(n) if bits is None else bits.shape[1]
        if bits is None:
            bits = np.zeros((self.channels, mb), dtype=np.uint8)
        if nbits is None:
            nbits = np.zeros(self.channels, dtype=np.int32)
        if tags is None:
            tags = np.zeros((self.channels, self.max_tags), dtype=B.TAG_DTYPE)
        if ntags is None:
            ntags = np.zeros(self.channels, dtype=np.int32)
        B.check(B.lib().b200ais_demod_work_sc16(self._h, B.ptr(iq16), C.c_float(scale), n, B.ptr(bits), mb,
                                                B.ptr(nbits), B.ptr(tags), B.ptr(ntags)))
        return bits, nbits, tags, ntags

    def work_dev(self, iq_ptr, nsamples, bits_ptr, max_bits, nbits_ptr, tags_ptr=None,
                 ntags_ptr=None, stream=None):
        """Device-resident variant: raw device addresses (e.g. torch_tensor.data_ptr()),
        asynchronous on `stream` (a cudaStream_t value; None = the default stream)."""
        B.check(B.lib().b200ais_demod_work_dev(self._h, B.ptr(iq_ptr), int(nsamples), B.ptr(bits_ptr),
                                               int(max_bits), B.ptr(nbits_ptr), B.ptr(tags_ptr),
                                               B.ptr(ntags_ptr), stream))

    def set_symbols(self, symbols, stream=None):
        """corr_est_cc::set_symbols on the chain's correlator: taps replaced verbatim, threshold
        unchanged (lib/corr_est_cc_impl.cc:132-162); same length as at construction."""
        symbols = np.ascontiguousarray(symbols, dtype=np.complex64)
        B.check(B.lib().b200ais_demod_set_symbols(self._h, B.ptr(symbols), len(symbols), stream))
        self.mod_vector = symbols

    def enqueue_dev(self, iq_ptr, nsamples, bits_ptr, max_bits, nbits_ptr, tags_ptr=None,
                    ntags_ptr=None, stream=None):
        """work_dev for a run of independent records: the timing loop of this record runs on
        an internal high-priority stream under the front half of the next one.  bits / nbits
        are complete after join(stream)."""
        B.check(B.lib().b200ais_demod_enqueue_dev(self._h, B.ptr(iq_ptr), int(nsamples), B.ptr(bits_ptr),
                                                  int(max_bits), B.ptr(nbits_ptr), B.ptr(tags_ptr),
                                                  B.ptr(ntags_ptr), stream))

    def join(self, stream=None):
        B.check(B.lib().b200ais_demod_join(self._h, stream))

    # ---- the chain as a stream: the blocks keep their state from call to call ----

    def stream_reset(self, stream=None):
        """Back to freshly constructed blocks (also implied by the first stream call)."""
        B.check(B.lib().b200ais_demod_stream_reset(self._h, stream))

    def stream_max_bits(self, nsamples):
        return B.lib().b200ais_demod_stream_max_bits(self._h, int(nsamples))

    def stream_work(self, iq):
        """Feed the next piece of every channel's capture: iq [channels, n] complex64 host
        array, any n in [0, max_samples].  Returns this call's (bits, nbits, tags, ntags);
        tag offsets are absolute corr_est item offsets."""
        iq = np.asarray(iq)
        if iq.dtype != np.complex64 or not iq.flags.c_contiguous:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim == 1:
            iq = iq.reshape(1, -1)
        if iq.shape[0] != self.channels:
            raise ValueError("expected %d channels" % self.channels)
        n = iq.shape[1]
        mb = self.stream_max_bits(n)
        bits = np.zeros((self.channels, mb), dtype=np.uint8)
        nbits = np.zeros(self.channels, dtype=np.int32)
        tags = np.zeros((self.channels, self.max_tags), dtype=B.TAG_DTYPE)
        ntags = np.zeros(self.channels, dtype=np.int32)
        B.check(B.lib().b200ais_demod_stream_work(self._h, B.ptr(iq), n, B.ptr(bits), mb, B.ptr(nbits),
                                                  B.ptr(tags), B.ptr(ntags)))
        return bits, nbits, tags, ntags

    def stream_work_dev(self, iq_ptr, nsamples, bits_ptr, max_bits, nbits_ptr, tags_ptr=None,
                        ntags_ptr=None, stream=None):
        """Device-resident variant of stream_work (raw device addresses, asynchronous)."""
        B.check(B.lib().b200ais_demod_stream_work_dev(self._h, B.ptr(iq_ptr), int(nsamples),
                                                      B.ptr(bits_ptr), int(max_bits), B.ptr(nbits_ptr),
                                                      B.ptr(tags_ptr), B.ptr(ntags_ptr), stream))

    def stream_pending(self):
        """(input items short of an FFT vector, AGC outputs short of a corr_est output
        multiple, corr_est nitems_written)."""
        a, b, w = C.c_int(0), C.c_int(0), C.c_uint64(0)
        B.check(B.lib().b200ais_demod_stream_pending(self._h, C.byref(a), C.byref(b), C.byref(w)))
        return a.value, b.value, w.value

    def set_overlap(self, groups):
        B.check(B.lib().b200ais_demod_set_overlap(self._h, int(groups)))

    def profile(self, on=True):
        B.check(B.lib().b200ais_demod_profile(self._h, 1 if on else 0))

    def stage_ms(self):
        """(dict stage -> summed ms, calls) since profiling was enabled / last read."""
        ms = (C.c_double * len(B.STAGE_NAMES))()
        calls = C.c_int(0)
        B.check(B.lib().b200ais_demod_stage_ms(self._h, ms, C.byref(calls)))
        return dict(zip(B.STAGE_NAMES, list(ms))), calls.value

    def status(self):
        B.check(B.lib().b200ais_demod_status(self._h))

    def read_tap(self, which, nsamples=None):
        """Copy one intermediate stream of the last work call to the host."""
        p, row = C.c_void_p(), C.c_size_t()
        B.check(B.lib().b200ais_demod_tap(self._h, which, C.byref(p), C.byref(row)))
        dt = {B.TAP_FHAT: np.float32, B.TAP_AGC: np.complex64, B.TAP_SYM: np.complex64,
              B.TAP_ERR: np.float32, B.TAP_MU: np.float32, B.TAP_SOFT: np.float32,
              B.TAP_MASK: np.uint8}[which]
        out = np.zeros((self.channels, max(row.value, 1)), dtype=dt)
        B.check(B.lib().b200ais_demod_read_tap(self._h, which, B.ptr(out), out.nbytes))
        return out[:, :row.value]


class square_and_fft_sync_cc:
    """x -> x * exp(-j 2 pi fhat n / fs) with fhat from squaring + FFT + ais.freqest
    (python/gmsk_sync.py:14-37), as a stand-alone stage: the chain with AGC switched off,
    read at the corr_est input."""

    def __init__(self, samplerate, bits_per_sec, fftlen, channels=1, max_samples=48000):
        opts = dict(default_options(), bits_per_sec=float(bits_per_sec), fftlen=int(fftlen),
                    samples_per_symbol=float(samplerate) / float(bits_per_sec))
        self._d = ais_demod(opts, channels=channels, max_samples=max_samples,
                            stages=B.STAGE_FREQSYNC)

    def work(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim == 1:
            iq = iq.reshape(1, -1)
        self._d.work(iq)
        n1 = (iq.shape[1] // self._d.fftlen) * self._d.fftlen
        return self._d.read_tap(B.TAP_AGC)[:, :n1].copy(), self._d.read_tap(B.TAP_FHAT).copy()
