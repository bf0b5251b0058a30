"""CPU-only checks of the host side: the C-ABI library loads and exports what the header
declares, the host mirrors behave like the reference's interface, the synthetic-traffic
framing is self-consistent, and nothing silently falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from gr_ais_b200 import binding as B
from gr_ais_b200 import sharding, synth
from gr_ais_b200 import ais_demod as AD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_present():
    n = ctypes.c_int(0)
    return B.lib().b200ais_device_count(ctypes.byref(n)) == 0 and n.value > 0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200ais.h")).read()
    declared = sorted(set(re.findall(r"B200AIS_API[^;(]*?\b(b200ais_\w+)\s*\(", hdr)))
    assert declared == sorted(B.EXPORTS)
    lib = B.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.b200ais_version() >= 100


def test_header_cites_the_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "b200ais.h")).read()
    for cite in ("lib/corr_est_cc_impl.cc", "lib/msk_timing_recovery_cc_impl.cc",
                 "lib/freqest_impl.cc", "lib/invert_impl.cc", "python/ais_demod.py",
                 "python/gmsk_sync.py"):
        assert cite in hdr


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU every compute entry point must fail loudly (B200AIS_E_CUDA)."""
    if _cuda_present():
        pytest.skip("a CUDA device is present")
    tmpl = np.ones(16, np.complex64)
    with pytest.raises(B.B200AisError) as e:
        AD.ais_demod(channels=1, max_samples=2048, template=tmpl)
    assert e.value.code in (B.E_CUDA, B.E_NOMEM)
    from gr_ais_b200 import blocks
    with pytest.raises(B.B200AisError):
        blocks.corr_est_cc.make(tmpl, 5.0, 1, 0.9)
    with pytest.raises(B.B200AisError):
        blocks.freqest.make(48000.0, 9600, 1024).work(1, [np.zeros(1024, np.complex64)],
                                                      [np.zeros(1, np.float32)])
    out = np.zeros(4, np.uint8)
    with pytest.raises(B.B200AisError):
        blocks.invert.make().work(4, [np.zeros(4, np.uint8)], [out])


def test_reference_error_behaviour_needs_no_device():
    """Argument errors the reference raises in its constructors surface before any CUDA call."""
    from gr_ais_b200 import blocks
    with pytest.raises(IndexError, match="Gain must be positive"):
        blocks.msk_timing_recovery_cc.make(5.0, 0.0, 0.01, 1)
    with pytest.raises(IndexError, match="osps must be 1 or 2"):
        blocks.msk_timing_recovery_cc.make(5.0, 0.04, 0.01, 3)
    with pytest.raises(IndexError):
        AD.ais_demod({"clockrec_gain": -1.0}, template=np.ones(8, np.complex64))


def test_default_config_matches_the_reference_options():
    cfg = B.default_config()
    # python/radio.py:47-62, python/ais_demod.py:35,41-42
    assert (cfg.sample_rate, cfg.data_rate, cfg.fftlen) == (48000.0, 9600, 1024)
    assert (cfg.agc_nsamples, cfg.agc_reference) == (512, 2.0)
    assert cfg.sps == 5.0 and cfg.mark_delay == 1 and cfg.osps == 1
    assert cfg.threshold == pytest.approx(0.9) and cfg.gain == pytest.approx(0.04)
    assert cfg.limit == pytest.approx(0.01)
    assert cfg.stages == B.STAGE_FREQSYNC | B.STAGE_AGC
    opts = AD.default_options()
    assert opts["fftlen"] == 1024 and opts["clockrec_gain"] == 0.04


def test_template_mirror_equals_oracle(oracle):
    """Two independent restatements of gmsk_mod + modulate_vector_bc agree bit for bit."""
    assert np.array_equal(AD.preamble_template("north_star"),
                          oracle.gmsk_template_bits(np.array([1, 1, 0, 0] * 6, np.uint8)))
    assert np.array_equal(AD.preamble_template("intended"),
                          oracle.gmsk_template_bits(np.array([1, 1, 0, 0] * 7, np.uint8)))
    ref = AD.preamble_template("reference")
    assert len(ref) == 1120   # 28 packed bytes -> 224 bits x 5 samples (python/ais_demod.py:36-38)
    assert np.array_equal(ref, oracle.gmsk_template_packed(np.array([1, 1, 0, 0] * 7, np.uint8)))
    assert np.allclose(np.abs(ref), 1.0, atol=1e-5)
    assert np.array_equal(AD.firdes_gaussian(1, 5, 0.4, 20), oracle.firdes_gaussian(1, 5, 0.4, 20))


def test_template_is_gmsk(oracle):
    t = AD.preamble_template("north_star")
    # 1,1,0,0 repeating: after the Gaussian filter's start-up transient the waveform repeats
    # every 4 symbols (20 samples), swings by less than pi/2 per symbol and has no net rotation
    ph = np.unwrap(np.angle(t))
    assert np.max(np.abs(t[40:100] - t[60:120])) < 1e-3
    per_symbol = ph[25::5][1:] - ph[25::5][:-1]
    assert np.all(np.abs(per_symbol) <= np.pi / 2 + 0.05)
    assert abs(ph[100] - ph[40]) < 1e-2


def test_crc_and_framing_round_trip():
    assert synth.crc16_x25(b"123456789") == 0x906E
    rng = np.random.default_rng(0)
    for _ in range(20):
        payload = synth.random_payload(rng)
        bits = synth.frame_bits(payload)
        assert synth.hdlc_deframe(bits) == [payload]
        # NRZI as the demod undoes it: bit = not(level xor previous level)
        lv = synth.nrzi_encode(bits, synth.nrzi_level_for_training())
        dec = 1 - (lv ^ np.concatenate([[synth.nrzi_level_for_training()], lv[:-1]]))
        assert np.array_equal(dec, bits)
    stuffed = synth.hdlc_stuff([1] * 12)
    assert list(stuffed) == [1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 0, 1, 1]


def test_synthetic_records_are_reproducible():
    a, ta = synth.make_record(3, n=8192, nbursts=2)
    b, tb = synth.make_record(3, n=8192, nbursts=2)
    assert np.array_equal(a, b) and ta == tb and a.dtype == np.complex64
    c, _ = synth.make_record(4, n=8192, nbursts=2)
    assert not np.array_equal(a, c)
    r = synth.replicate_record(a, 3)
    assert np.array_equal(r[2], np.roll(a, 32))


def test_oracle_chain_decodes_known_payloads(oracle, templates):
    """KAT 1: payload in => payload out, through the CPU oracle."""
    found = total = 0
    for ch in range(4):
        x, truth = synth.make_record(ch, n=48000, nbursts=4, snr_db=25)
        r = oracle.demod_chain(x, templates[120], oracle.chain_cfg(threshold=2.0))
        f = synth.payloads_found(r["bits"], truth)
        found += sum(f)
        total += len(f)
    assert total == 16 and found >= 15


def test_oracle_chain_scheduling_contract(oracle, templates):
    x, _ = synth.make_record(0, n=5000, nbursts=1)
    r = oracle.demod_chain(x, templates[120], debug=True)
    assert r["n1"] == 4096                      # whole FFT vectors only (stream_to_vector)
    assert r["n2"] == (4096 // 137) * 137       # corr_est output multiple (fft_filter nsamples)
    assert oracle.default_corr_chunk(120) == (24576 // 137) * 137
    assert len(r["fhat"]) == 4 and len(r["bits"]) == len(r["sym"])
    # corr_est delays by L: msk sees zeros for the first L items
    assert np.all(r["agc"][:10] == 0)           # AGC history is 511 zeros too


def test_partition_covers_every_channel_once():
    for C, W in ((4096, 8), (10, 3), (7, 8), (262144, 8)):
        spans = [sharding.partition(C, W, r) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == C
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.partition(8, 2, 2)


def test_broadcast_template_single_process():
    t = np.arange(6).astype(np.complex64)
    assert np.array_equal(sharding.broadcast_template(t), t)
    assert sharding.max_over_ranks(1.5) == 1.5


def test_numa_binding_is_a_no_op_without_nvml_devices():
    import os
    from gr_ais_b200 import sharding
    before = os.sched_getaffinity(0)
    assert sharding.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/b200ais.h must compile as C99 on its own."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "b200ais.h"\nint main(void) { b200ais_rx_config c; (void)c; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(root, "include"), str(src)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_receiver_rows_have_no_cpu_fallback_either(oracle):
    """firdes.low_pass is init-time host code (as in the reference) and works anywhere; the
    channeliser, deframer, NMEA formatter and ais_rx need the device and say so."""
    from gr_ais_b200 import blocks
    from gr_ais_b200.radio import ais_rx
    taps = blocks.firdes_low_pass(1.0, 250e3, 11e3, 1e3)
    assert np.array_equal(taps, oracle.firdes_low_pass(1.0, 250e3, 11e3, 1e3))
    with pytest.raises(B.B200AisError) as e:
        blocks.firdes_low_pass(1.0, 250e3, 200e3, 1e3)     # cutoff beyond fs/2: std::out_of_range
    assert e.value.code == B.E_RANGE
    if _cuda_present():
        pytest.skip("a CUDA device is present")
    for make in (lambda: blocks.freq_xlating_fir_filter_ccf(5, taps, [-25e3, 25e3], 250e3),
                 lambda: blocks.hdlc_deframer_bp(11, 64),
                 lambda: ais_rx([-25e3, 25e3], 250e3, ["A", "B"])):
        with pytest.raises(B.B200AisError) as e:
            make()
        assert e.value.code in (B.E_CUDA, B.E_NOMEM)
    with pytest.raises(B.B200AisError):
        blocks.pdu_to_nmea("A").to_nmea(bytes(21))


def test_rx_default_config_matches_the_reference_receiver():
    import ctypes as C
    cfg = B.RxConfig()
    B.check(B.lib().b200ais_rx_default_config(C.byref(cfg)))
    assert cfg.rate == 250e3 and cfg.nfreqs == 2                  # python/radio.py:120, :88-89
    assert (cfg.freqs[0], cfg.freqs[1]) == (161.975e6 - 162.0e6, 162.025e6 - 162.0e6)
    assert cfg.designators[0].value == b"A" and cfg.designators[1].value == b"B"
    assert (cfg.hdlc_length_min, cfg.hdlc_length_max) == (11, 64)  # :64
    assert (cfg.lpf_cutoff, cfg.lpf_transition) == (11000.0, 1000.0)  # :49
    assert abs(cfg.clockrec_gain - 0.04) < 1e-7 and abs(cfg.omega_relative_limit - 0.01) < 1e-7
    assert cfg.fftlen == 1024 and cfg.bits_per_sec == 9600.0
