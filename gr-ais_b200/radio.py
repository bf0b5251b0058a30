"""Host-side mirror of the reference's receiver hier-block ais_rx (python/radio.py:39-72) on top
of the C-ABI (b200ais_rx_*): wideband complex IQ in, NMEA AIVDM sentences out.

    freq_xlating_fir_filter_ccf(int(rate/48000), firdes.low_pass(1, rate, 11000, 1000), freq, rate)
    -> ais_demod(options) -> hdlc_deframer_bp(11, 64) -> pdu_to_nmea(designator)

The reference instantiates one ais_rx per AIS channel (A at -25 kHz, B at +25 kHz) on the same
source (python/radio.py:86-91); here one object carries every (source, frequency) pair:
`freq` / `designator` may be lists, `sources` wideband inputs are processed side by side.
Fails loudly without the CUDA library or a device: there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import binding as B
from .ais_demod import preamble_template


class ais_rx:
    def __init__(self, freq, rate, designator, sources=1, max_input_items=1 << 18, max_frames=64,
                 template="north_star", options=None):
        freqs = [float(f) for f in np.atleast_1d(freq)]
        des = [designator] if isinstance(designator, str) else list(designator)
        if len(des) != len(freqs):
            raise ValueError("one designator per frequency")
        cfg = B.RxConfig()
        B.check(B.lib().b200ais_rx_default_config(C.byref(cfg)))
        cfg.rate = float(rate)
        cfg.nfreqs = len(freqs)
        for k, f in enumerate(freqs):
            cfg.freqs[k] = f
            d = des[k].encode()
            if not 0 < len(d) <= 8:
                raise ValueError("designator: 1..8 characters")
            cfg.designators[k].value = d
        cfg.sources = int(sources)
        cfg.max_input_items = int(max_input_items)
        cfg.max_frames = int(max_frames)
        for k, v in (options or {}).items():  # clockrec_gain, omega_relative_limit, fftlen, ...
            setattr(cfg, k, v)
        self.cfg = cfg
        self.freqs, self.designators = freqs, des
        self.sources, self.rate = int(sources), float(rate)
        # python/radio.py:50,57: integer decimation, the demod runs at whatever rate results
        self._filter_decimation = int(self.rate / (cfg.bits_per_sec * 5))
        self.samples_per_symbol = (self.rate / self._filter_decimation) / cfg.bits_per_sec
        if isinstance(template, str):
            # digital.gmsk_mod takes an integer samples_per_symbol (python/ais_demod.py:37)
            template = preamble_template(template, int(self.samples_per_symbol))
        self.mod_vector = np.ascontiguousarray(template, dtype=np.complex64)
        self._h = C.c_void_p()
        B.check(B.lib().b200ais_rx_create(C.byref(self._h), C.byref(cfg), B.ptr(self.mod_vector),
                                          len(self.mod_vector)))
        self.channels = B.lib().b200ais_rx_channels(self._h)
        self.slot = B.lib().b200ais_rx_sentence_slot(self._h)

    def __del__(self):
        self.close()

    def close(self):
        try:
            if self._h:
                B.lib().b200ais_rx_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def reset(self):
        B.check(B.lib().b200ais_rx_reset(self._h))

    def decimation(self):
        return B.lib().b200ais_rx_decimation(self._h)

    def work(self, iq, max_msgs=None):
        """iq: [sources, n] complex64 host array -- the next n items of every source.
        Returns (msgs, sentences): msgs a structured array (binding.FRAME_DTYPE; channel =
        source*len(freq) + k) sorted by (channel, end_bit), sentences the matching strings."""
        iq = np.asarray(iq)
        if iq.dtype != np.complex64 or not iq.flags.c_contiguous:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim == 1:
            iq = iq.reshape(1, -1)
        if iq.shape[0] != self.sources:
            raise ValueError("expected %d sources" % self.sources)
        if max_msgs is None:
            max_msgs = self.channels * self.cfg.max_frames
        msgs = np.zeros(max_msgs, dtype=B.FRAME_DTYPE)
        sent = np.zeros((max_msgs, self.slot), dtype=np.uint8)
        lens = np.zeros(max_msgs, dtype=np.int32)
        n = C.c_int(0)
        B.check(B.lib().b200ais_rx_work(self._h, B.ptr(iq), iq.shape[1], iq.shape[1], B.ptr(msgs),
                                        B.ptr(sent), self.slot, B.ptr(lens), max_msgs, C.byref(n)))
        k = n.value
        order = np.lexsort((msgs["end_bit"][:k], msgs["channel"][:k]))
        return msgs[:k][order], [bytes(sent[i, :lens[i]]).decode("latin-1") for i in order]

    def work_dev(self, iq_ptr, iq_stride, nitems, msgs_ptr, sent_ptr, lens_ptr, max_msgs, nmsgs_ptr,
                 stream=None):
        """Device-resident, asynchronous variant (raw device addresses)."""
        B.check(B.lib().b200ais_rx_work_dev(self._h, B.ptr(iq_ptr), int(iq_stride), int(nitems),
                                            B.ptr(msgs_ptr), B.ptr(sent_ptr), self.slot,
                                            B.ptr(lens_ptr), int(max_msgs), B.ptr(nmsgs_ptr), stream))

    def replay_file(self, path, chunk_items=None, max_msgs=None):
        """blocks.file_source(gr.sizeof_gr_complex, path) into every source (python/radio.py:
        211-213): raw interleaved float32 IQ, double-buffered pinned reads.  Returns
        (msgs, sentences, items_read) over the whole file, sorted by (channel, end_bit)."""
        return self._pump(lambda cb, items, chunk, mm: B.lib().b200ais_rx_replay_file(
            self._h, os.fsencode(path), chunk, mm, cb, None, C.byref(items)), chunk_items, max_msgs)

    def serve_udp(self, ip, port, chunk_items=None, max_msgs=None, max_items=0, idle_ms=1000):
        """blocks.udp_source(gr.sizeof_gr_complex, ip, port) into every source (python/radio.py:
        204-210).  Blocks until a zero-length datagram, max_items items or idle_ms of silence."""
        return self._pump(lambda cb, items, chunk, mm: B.lib().b200ais_rx_serve_udp(
            self._h, ip.encode(), int(port), chunk, mm, int(max_items), int(idle_ms), cb, None,
            C.byref(items)), chunk_items, max_msgs)

    def _pump(self, call, chunk_items, max_msgs):
        chunk_items = int(chunk_items or self.cfg.max_input_items)
        max_msgs = int(max_msgs or self.channels * self.cfg.max_frames)
        got_m, got_s = [], []

        def sink(user, msgs, sent, slot, lens, n):
            m = np.frombuffer((C.c_char * (n * B.FRAME_DTYPE.itemsize)).from_address(msgs),
                              dtype=B.FRAME_DTYPE).copy()
            ln = np.frombuffer((C.c_char * (4 * n)).from_address(lens), dtype=np.int32)
            raw = np.frombuffer((C.c_char * (n * slot)).from_address(sent), dtype=np.uint8).reshape(n, slot)
            got_m.append(m)
            got_s.extend(bytes(raw[i, :ln[i]]).decode("latin-1") for i in range(n))

        cb = B.RX_SINK(sink)
        items = C.c_uint64(0)
        B.check(call(cb, items, chunk_items, max_msgs))
        msgs = np.concatenate(got_m) if got_m else np.zeros(0, dtype=B.FRAME_DTYPE)
        order = np.lexsort((msgs["end_bit"], msgs["channel"]))
        return msgs[order], [got_s[i] for i in order], items.value

    def tag_overflows(self):
        """calls in which a channel met more corr_est tags than its row holds (extra tags dropped)"""
        return int(B.lib().b200ais_rx_tag_overflows(self._h))

    def status(self):
        B.check(B.lib().b200ais_rx_status(self._h))
