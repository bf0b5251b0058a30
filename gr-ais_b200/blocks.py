"""Host-side mirrors of the reference's four C++ blocks, on top of the C-ABI.

Same names, factory arguments, accessors, work()/general_work() argument meaning
and error behaviour as gr::ais::{corr_est_cc, msk_timing_recovery_cc, freqest,
invert} (reference include/ais/*.h, lib/*_impl.cc), so tests read like GNU Radio
QA code.  Each block can carry `channels` independent streams (the reference
block is the channels = 1 case): arrays are then [channels, items].

GNU Radio is not available in this environment, so the scheduler-side services a
block relies on are modelled minimally here: history(), output_multiple(),
nitems_written()/nitems_read() counters and a per-call list of added stream tags.
The C++ adapter with the real gr::block signatures is gr-ais_b200/csrc/gr_adapter.
"""
import ctypes as C

import numpy as np

from . import binding as B

TAG_KEYS = {B.TAG_CORR_START: "corr_start", B.TAG_PHASE_EST: "phase_est",
            B.TAG_TIME_EST: "time_est", B.TAG_CORR_EST: "corr_est"}


def _as_c64(a, channels):
    a = np.ascontiguousarray(a, dtype=np.complex64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape[0] != channels:
        raise ValueError("expected %d channel rows, got %d" % (channels, a.shape[0]))
    return a


class corr_est_cc:
    """gr::ais::corr_est_cc (include/ais/corr_est_cc.h:85-107)."""

    def __init__(self, symbols, sps, mark_delay, threshold=0.9, channels=1):
        symbols = np.ascontiguousarray(symbols, dtype=np.complex64)
        self._h = C.c_void_p()
        self.channels = int(channels)
        B.check(B.lib().b200ais_corr_est_create(C.byref(self._h), B.ptr(symbols), len(symbols),
                                                float(sps), int(mark_delay), float(threshold),
                                                self.channels))
        self._written = 0
        self.tags = []  # structured array per work() call, one list entry per channel

    @classmethod
    def make(cls, symbols, sps, mark_delay, threshold=0.9, channels=1):
        return cls(symbols, sps, mark_delay, threshold, channels)

    def __del__(self):
        try:
            if self._h:
                B.lib().b200ais_corr_est_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # -- accessors the reference exposes
    def symbols(self):
        n = C.c_int(0)
        B.check(B.lib().b200ais_corr_est_symbols(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.complex64)
        B.check(B.lib().b200ais_corr_est_symbols(self._h, B.ptr(out), n.value, C.byref(n)))
        return out

    def set_symbols(self, symbols):
        symbols = np.ascontiguousarray(symbols, dtype=np.complex64)
        B.check(B.lib().b200ais_corr_est_set_symbols(self._h, B.ptr(symbols), len(symbols)))

    # -- scheduler-facing properties (lib/corr_est_cc_impl.cc:85,95-98,112)
    def history(self):
        return B.lib().b200ais_corr_est_history(self._h)

    def output_multiple(self):
        return B.lib().b200ais_corr_est_output_multiple(self._h)

    def max_noutput_items(self):
        return 24 * 1024

    def mark_delay(self):
        return B.lib().b200ais_corr_est_mark_delay(self._h)

    def threshold(self):
        return B.lib().b200ais_corr_est_threshold(self._h)

    def nitems_written(self, port=0):
        return self._written

    def work(self, noutput_items, input_items, output_items, max_tags=1024):
        """input_items[0]: noutput_items + history()-1 items per channel (history first);
        output_items: [out0] or [out0, out1].  Returns noutput_items; tags land in self.tags."""
        n = int(noutput_items)
        L = self.history() - 1
        inp = _as_c64(input_items[0], self.channels)
        if inp.shape[1] < n + L:
            raise ValueError("input holds %d items, work() needs %d" % (inp.shape[1], n + L))
        out0 = output_items[0] if len(output_items) > 0 else None
        out1 = output_items[1] if len(output_items) > 1 else None
        for o in (out0, out1):
            if o is not None and (o.dtype != np.complex64 or not o.flags.c_contiguous):
                raise ValueError("outputs must be C-contiguous complex64")
        ostride = 0
        for o in (out0, out1):
            if o is not None:
                ostride = o.shape[-1]
        tags = np.zeros((self.channels, max_tags), dtype=B.TAG_DTYPE)
        ntags = np.zeros(self.channels, dtype=np.int32)
        B.check(B.lib().b200ais_corr_est_work(self._h, n, B.ptr(inp), inp.shape[1], self._written,
                                              B.ptr(out0), B.ptr(out1), ostride, B.ptr(tags),
                                              max_tags, B.ptr(ntags)))
        self.tags = [tags[c, :ntags[c]].copy() for c in range(self.channels)]
        self._written += n
        return n


class msk_timing_recovery_cc:
    """gr::ais::msk_timing_recovery_cc (include/ais/msk_timing_recovery_cc.h:46-70)."""

    def __init__(self, sps, gain, limit, osps=1, channels=1):
        self._h = C.c_void_p()
        self.channels = int(channels)
        rc = B.lib().b200ais_msk_create(C.byref(self._h), float(sps), float(gain), float(limit),
                                        int(osps), self.channels)
        if rc == B.E_RANGE:  # std::out_of_range in the reference (:61,82)
            raise IndexError(B.lib().b200ais_last_error().decode())
        B.check(rc)
        self._read = 0
        self.consumed = np.zeros(self.channels, dtype=np.int32)

    @classmethod
    def make(cls, sps, gain, limit, osps=1, channels=1):
        return cls(sps, gain, limit, osps, channels)

    def __del__(self):
        try:
            if self._h:
                B.lib().b200ais_msk_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def set_gain(self, gain):
        rc = B.lib().b200ais_msk_set_gain(self._h, float(gain))
        if rc == B.E_RANGE:
            raise IndexError(B.lib().b200ais_last_error().decode())
        B.check(rc)

    def get_gain(self):
        return B.lib().b200ais_msk_get_gain(self._h)

    def set_limit(self, limit):
        B.check(B.lib().b200ais_msk_set_limit(self._h, float(limit)))

    def get_limit(self):
        return B.lib().b200ais_msk_get_limit(self._h)

    def set_sps(self, sps):
        B.check(B.lib().b200ais_msk_set_sps(self._h, float(sps)))

    def get_sps(self):
        return B.lib().b200ais_msk_get_sps(self._h)

    def forecast(self, noutput_items, ninput_items_required=None):
        need = B.lib().b200ais_msk_forecast(self._h, int(noutput_items))
        if ninput_items_required is not None:
            for i in range(len(ninput_items_required)):
                ninput_items_required[i] = need
        return need

    def nitems_read(self, port=0):
        return self._read

    def reset(self):
        B.check(B.lib().b200ais_msk_reset(self._h))
        self._read = 0

    def general_work(self, noutput_items, ninput_items, input_items, output_items, tags=None):
        """tags: per-channel list of structured arrays (B.TAG_DTYPE) visible on the input, as
        get_tags_in_range would return them.  Returns items produced per channel (an int when
        channels == 1); self.consumed holds the consume_each() arguments."""
        nin = int(ninput_items[0]) if np.ndim(ninput_items) else int(ninput_items)
        nout = int(noutput_items)
        inp = _as_c64(input_items[0], self.channels)
        if inp.shape[1] < nin:
            raise ValueError("input holds fewer than ninput_items items")
        out = output_items[0]
        err = output_items[1] if len(output_items) > 1 else None
        mu = output_items[2] if len(output_items) > 2 else None
        ostride = out.shape[-1]
        mt = 0
        tg = nt = None
        if tags is not None:
            if isinstance(tags, np.ndarray) and tags.ndim == 1:
                tags = [tags]
            mt = max(1, max(len(t) for t in tags))
            tg = np.zeros((self.channels, mt), dtype=B.TAG_DTYPE)
            nt = np.zeros(self.channels, dtype=np.int32)
            for c, t in enumerate(tags):
                tg[c, :len(t)] = t
                nt[c] = len(t)
        nprod = np.zeros(self.channels, dtype=np.int32)
        ncons = np.zeros(self.channels, dtype=np.int32)
        B.check(B.lib().b200ais_msk_general_work(self._h, nout, nin, B.ptr(inp), inp.shape[1],
                                                 self._read, B.ptr(tg), mt, B.ptr(nt), B.ptr(out),
                                                 B.ptr(err), B.ptr(mu), ostride, B.ptr(nprod),
                                                 B.ptr(ncons)))
        self.consumed = ncons
        if self.channels == 1:
            self._read += int(ncons[0])
            return int(nprod[0])
        return nprod


class freqest:
    """gr::ais::freqest (include/ais/freqest.h:37-50)."""

    def __init__(self, sample_rate, data_rate, fftlen, channels=1):
        self._h = C.c_void_p()
        self.channels = int(channels)
        self.fftlen = int(fftlen)
        B.check(B.lib().b200ais_freqest_create(C.byref(self._h), float(sample_rate), int(data_rate),
                                               int(fftlen), self.channels))

    @classmethod
    def make(cls, sample_rate, data_rate, fftlen, channels=1):
        return cls(sample_rate, data_rate, fftlen, channels)

    def __del__(self):
        try:
            if self._h:
                B.lib().b200ais_freqest_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def work(self, noutput_items, input_items, output_items):
        n = int(noutput_items)
        spec = np.ascontiguousarray(input_items[0], dtype=np.complex64).reshape(self.channels, -1)
        if spec.shape[1] < n * self.fftlen:
            raise ValueError("input holds fewer than noutput_items vectors")
        spec = np.ascontiguousarray(spec[:, :n * self.fftlen])
        out = output_items[0]
        tmp = np.zeros((self.channels, n), dtype=np.float32)
        B.check(B.lib().b200ais_freqest_work(self._h, n, B.ptr(spec), B.ptr(tmp)))
        out.reshape(self.channels, -1)[:, :n] = tmp
        return n


class invert:
    """gr::ais::invert (include/ais/invert.h:37-50)."""

    @classmethod
    def make(cls):
        return cls()

    def work(self, noutput_items, input_items, output_items):
        n = int(noutput_items)
        inp = np.ascontiguousarray(input_items[0], dtype=np.uint8).reshape(-1)[:n]
        tmp = np.zeros(n, dtype=np.uint8)
        B.check(B.lib().b200ais_invert_work(B.ptr(inp), B.ptr(tmp), n))
        output_items[0].reshape(-1)[:n] = tmp
        return n


# ------------------------------------------------------------------------------------
# The blocks either side of ais_demod inside the reference's ais_rx (python/radio.py:39-72)

def firdes_low_pass(gain, sampling_freq, cutoff_freq, transition_width):
    """filter.firdes.low_pass (Hamming window), as python/radio.py:49 calls it."""
    n = C.c_int(0)
    B.check(B.lib().b200ais_firdes_low_pass(gain, sampling_freq, cutoff_freq, transition_width,
                                            None, 0, C.byref(n)))
    taps = np.zeros(n.value, dtype=np.float32)
    B.check(B.lib().b200ais_firdes_low_pass(gain, sampling_freq, cutoff_freq, transition_width,
                                            B.ptr(taps), n.value, C.byref(n)))
    return taps


class freq_xlating_fir_filter_ccf:
    """filter.freq_xlating_fir_filter_ccf(decimation, taps, center_freq, sampling_freq)
    (python/radio.py:51-54).  center_freq may be a list: that many filters share each of the
    `sources` inputs (the A and B rx paths of python/radio.py:86-91); output row
    s*len(center_freq) + k is source s translated by center_freq[k]."""

    def __init__(self, decimation, taps, center_freq, sampling_freq, sources=1):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        freqs = np.atleast_1d(np.asarray(center_freq, dtype=np.float64)).copy()
        self._h = C.c_void_p()
        self.sources, self.nfreqs, self.ntaps = int(sources), len(freqs), len(taps)
        self._decim = int(decimation)
        B.check(B.lib().b200ais_xlat_create(C.byref(self._h), self._decim, B.ptr(taps), len(taps),
                                            B.ptr(freqs), len(freqs), float(sampling_freq),
                                            self.sources))

    def __del__(self):
        try:
            if self._h:
                B.lib().b200ais_xlat_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def history(self):
        return B.lib().b200ais_xlat_history(self._h)

    def decimation(self):
        return B.lib().b200ais_xlat_decimation(self._h)

    def set_center_freq(self, center_freq, k=0):
        B.check(B.lib().b200ais_xlat_set_center_freq(self._h, int(k), float(center_freq)))

    def set_taps(self, taps):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        B.check(B.lib().b200ais_xlat_set_taps(self._h, B.ptr(taps), len(taps)))
        self.ntaps = len(taps)

    def reset(self):
        B.check(B.lib().b200ais_xlat_reset(self._h))

    def work(self, noutput_items, input_items, output_items):
        """input_items[0]: [sources, >= history()-1 + noutput_items*decimation()] (history
        first); output_items[0]: [sources*nfreqs, >= noutput_items].  Returns noutput_items."""
        x = _as_c64(input_items[0], self.sources)
        need = self.ntaps - 1 + noutput_items * self._decim
        if x.shape[1] < need:
            raise ValueError("work: %d input items per row, need %d" % (x.shape[1], need))
        out = output_items[0]
        if out.dtype != np.complex64 or out.shape[0] != self.sources * self.nfreqs or \
                out.shape[1] < noutput_items or not out.flags.c_contiguous:
            raise ValueError("work: bad output array")
        B.check(B.lib().b200ais_xlat_work(self._h, int(noutput_items), B.ptr(x), x.shape[1],
                                          B.ptr(out), out.shape[1]))
        return noutput_items


class hdlc_deframer_bp:
    """digital.hdlc_deframer_bp(length_min, length_max) (python/radio.py:64): unpacked bits in,
    one PDU (bytes, CRC checked and removed) per good frame out."""

    def __init__(self, length_min=11, length_max=64, channels=1):
        self._h = C.c_void_p()
        self.channels = int(channels)
        B.check(B.lib().b200ais_hdlc_create(C.byref(self._h), int(length_min), int(length_max),
                                            self.channels))

    def __del__(self):
        try:
            if self._h:
                B.lib().b200ais_hdlc_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def reset(self):
        B.check(B.lib().b200ais_hdlc_reset(self._h))

    def work(self, bits, nbits=None, max_frames=64):
        """bits: [channels, n] unpacked 0/1 bytes; nbits: valid items per row (default: all).
        Returns (frames [channels, max_frames] of binding.FRAME_DTYPE, nframes [channels])."""
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        if bits.ndim == 1:
            bits = bits.reshape(1, -1)
        if bits.shape[0] != self.channels:
            raise ValueError("expected %d rows" % self.channels)
        frames = np.zeros((self.channels, max_frames), dtype=B.FRAME_DTYPE)
        nframes = np.zeros(self.channels, dtype=np.int32)
        nb = None if nbits is None else np.ascontiguousarray(nbits, dtype=np.int32)
        B.check(B.lib().b200ais_hdlc_work(self._h, B.ptr(bits), bits.shape[1], B.ptr(nb),
                                          bits.shape[1], B.ptr(frames), max_frames, B.ptr(nframes)))
        return frames, nframes

    @staticmethod
    def pdus(frames, nframes):
        """the published PDUs per channel, as bytes"""
        return [[bytes(f["data"][:f["len"]]) for f in frames[c, :nframes[c]]]
                for c in range(frames.shape[0])]


class pdu_to_nmea:
    """gr::ais::pdu_to_nmea(designator) (include/ais/pdu_to_nmea.h:37-54)."""

    def __init__(self, designator="A"):
        self.designator = str(designator)
        if not 0 < len(self.designator.encode()) <= 8:
            raise ValueError("designator: 1..8 characters")

    @classmethod
    def make(cls, designator):
        return cls(designator)

    def format(self, frames, nframes, designators=None):
        """msg_to_sentence for every frame of hdlc_deframer_bp.work(); returns a list (per
        channel) of lists of sentence strings.  designators: per-channel override."""
        frames = np.ascontiguousarray(frames)
        nframes = np.ascontiguousarray(nframes, dtype=np.int32)
        channels, max_frames = frames.shape
        des = np.zeros((channels, 8), dtype=np.uint8)
        for c in range(channels):
            d = (designators[c] if designators is not None else self.designator).encode()
            des[c, :len(d)] = np.frombuffer(d, dtype=np.uint8)
        max_len = int(frames["len"].max()) if frames.size else 1
        slot = B.lib().b200ais_nmea_slot_bytes(max(max_len, 1), b"12345678")
        B.check(min(slot, 0))
        sent = np.zeros((channels, max_frames, slot), dtype=np.uint8)
        lens = np.zeros((channels, max_frames), dtype=np.int32)
        B.check(B.lib().b200ais_nmea_format(B.ptr(frames), B.ptr(nframes), channels, max_frames,
                                            B.ptr(des), B.ptr(sent), slot, B.ptr(lens)))
        if (lens < 0).any():
            raise B.B200AisError(B.E_OUT_OVERFLOW, "a sentence did not fit its slot")
        return [[bytes(sent[c, f, :lens[c, f]]).decode("latin-1") for f in range(nframes[c])]
                for c in range(channels)]

    def to_nmea(self, pdu: bytes) -> str:
        """the to_nmea / print message handlers for one PDU"""
        fr = np.zeros((1, 1), dtype=B.FRAME_DTYPE)
        fr[0, 0]["len"] = len(pdu)
        fr[0, 0]["data"][:len(pdu)] = np.frombuffer(bytes(pdu), dtype=np.uint8)
        return self.format(fr, np.array([1], dtype=np.int32))[0][0]
