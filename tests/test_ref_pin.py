"""Pins the oracle's [R] restatements to the reference's OWN code.

oracle/_ref/libais_ref.so is /root/reference/lib/{corr_est_cc,msk_timing_recovery_cc,freqest,
invert,pdu_to_nmea}_impl.cc compiled UNMODIFIED (oracle/ref_build/Makefile) against a stub of
the GNU Radio runtime; only the GNU Radio / VOLK kernels underneath (fft_filter_ccc, the MMSE
interpolator, fast_atan2f, branchless_clip, magnitude-squared) are the oracle's [G]
restatements.  Every test drives the reference's class and the oracle's restatement with the
same calls and requires identical bits: outputs, tag offsets/keys/values, consumed counts,
scheduler hints, exceptions.  The quirks SURVEY.md section 8a lists each have a case.

The library is built here when /root/reference is present and travels prebuilt to the GPU box
(tests/test_gpu_ref.py compares the CUDA path with it there)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and no reference tree")

TAG = {"corr_start": 0, "phase_est": 1, "time_est": 2, "corr_est": 3}


def same_tags(got, want):
    assert len(got) == len(want)
    for f in ("offset", "key", "port", "value"):
        assert np.array_equal(got[f], want[f], equal_nan=(f == "value")), f


def stream(n, seed, burst_at=(), tmpl=None, amp=1.0, noise=0.05):
    rng = np.random.default_rng(seed)
    x = (noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    for p in burst_at:
        x[p:p + len(tmpl)] += (amp * tmpl).astype(np.complex64)
    return x


def mk_tags(items):
    t = np.zeros(len(items), dtype=R.TAG_DTYPE)
    for k, (off, key, val) in enumerate(items):
        t[k] = (off, key, 0, val)
    return t


# --------------------------------------------------------------- provenance

def test_reference_sources_are_compiled_where_they_lie_unmodified():
    listed = dict((os.path.basename(l.split()[1]), l.split()[0]) for l in R.sources_sha256().splitlines())
    assert sorted(listed) == ["corr_est_cc_impl.cc", "freqest_impl.cc", "invert_impl.cc",
                              "msk_timing_recovery_cc_impl.cc", "pdu_to_nmea_impl.cc"]
    ref_lib = os.path.join(R.REFERENCE_ROOT, "lib")
    if os.path.isdir(ref_lib):       # this container; the GPU box only has the prebuilt library
        for name, digest in listed.items():
            with open(os.path.join(ref_lib, name), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, name
    # and no reference source was copied into the repository
    root = os.path.dirname(HERE)
    for d, _, files in os.walk(root):
        if "/.git" in d or "gpurun_out" in d:
            continue
        assert not any(f.endswith("_impl.cc") for f in files), d


# ---------------------------------------------------------------- corr_est_cc

@pytest.mark.parametrize("L", [120, 140, 1120])
def test_corr_est_ctor_hints_and_work(oracle, templates, L):
    t = templates[L]
    ref = R.CorrEstBlock(t, 5.0, 1, 0.9)
    ora = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
    fft = 2 * (1 << int(np.ceil(np.log2(L))))
    assert ref.hints() == dict(history=L + 1, output_multiple=fft - L + 1, max_noutput_items=24576,
                               sample_delay0=L, sample_delay1=0)
    assert ora.nsamples == ref.nsamples
    assert np.array_equal(ref.symbols(), ora.symbols())      # reverse(conj(template)) (:59-63)
    n = ref.nsamples * 3
    x = stream(2 * n + L, 11, burst_at=[L + 50, n + 300, 2 * n - 40], tmpl=t)
    written = 0
    for call in range(2):        # the third burst's peak sits near the edge of a work() chunk
        inbuf = x[written:written + n + L]
        a = ref.work(n, inbuf, nitems_written=written, two_ports=True)
        b = ora.work(n, inbuf, nitems_written=written, two_ports=True)
        for u, v in zip(a[:3], b[:3]):
            assert np.array_equal(u, v)
        same_tags(a[3], b[3])
        assert len(a[3]) > 0
        written += n


def test_corr_est_single_output_has_no_port1_tags(oracle, templates):
    t = templates[120]
    n = 137 * 4
    x = stream(n + 120, 21, burst_at=[180], tmpl=t)
    a = R.CorrEstBlock(t, 5.0, 1, 0.9).work(n, x)
    b = oracle.CorrEstBlock(t, 5.0, 1, 0.9).work(n, x)
    same_tags(a[3], b[3])
    assert set(a[3]["port"]) == {0}


@pytest.mark.parametrize("thr,md", [(0.9, 1), (0.5, 0), (0.2, 7), (0.9, 500)])
def test_corr_est_threshold_mark_delay_and_chunk_edges(oracle, templates, thr, md):
    """d_thresh = threshold*corr*corr (:71-74); mark_delay clamps to L-1 (:65-66); a peak at
    i == 0 or i == n-1 gets center 0.0 (:219-227); the climb stops at n-1 (:202-204)."""
    t = templates[120]
    ns = 137
    # bursts placed so that correlation peaks land on the first and on the last item of a chunk
    x = stream(4 * ns + 120, 5, burst_at=[1, ns - 119 + 119, 2 * ns + 17, 4 * ns - 119], tmpl=t, amp=1.1)
    ref, ora = R.CorrEstBlock(t, 5.0, md, thr), oracle.CorrEstBlock(t, 5.0, md, thr)
    for start, n in ((0, ns), (ns, ns), (2 * ns, 2 * ns)):
        a = ref.work(n, x[start:start + n + 120], nitems_written=start, two_ports=True)
        b = ora.work(n, x[start:start + n + 120], nitems_written=start, two_ports=True)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        same_tags(a[3], b[3])


def test_corr_est_set_symbols_quirk(oracle, templates):
    """set_symbols stores the taps verbatim (no conj, no reverse), keeps d_thresh and re-clamps
    the *current* mark_delay (lib/corr_est_cc_impl.cc:132-162)"""
    ref, ora = R.CorrEstBlock(templates[140], 5.0, 130, 0.9), oracle.CorrEstBlock(templates[140], 5.0, 130, 0.9)
    new = np.conj(templates[120])[::-1].copy()
    ref.set_symbols(new)
    ora.set_symbols(new)
    assert np.array_equal(ref.symbols(), new) and np.array_equal(ora.symbols(), new)
    assert ref.hints()["history"] == 121 and ref.nsamples == ora.nsamples == 137
    n = 137 * 3
    x = stream(n + 120, 6, burst_at=[200], tmpl=templates[120], amp=1.3)
    a, b = ref.work(n, x, two_ports=True), ora.work(n, x, two_ports=True)
    assert np.array_equal(a[1], b[1])
    same_tags(a[3], b[3])
    # threshold still that of the 140-tap template: 0.9 * 140^2 -- a unit burst of 120 taps stays below
    y = stream(n + 120, 7, burst_at=[200], tmpl=templates[120], amp=1.0, noise=0.0)
    assert len(ref.work(n, y)[3]) == len(ora.work(n, y)[3]) == 0


def test_corr_est_tail_carries_across_calls_and_template_swaps(oracle, templates):
    ref, ora = R.CorrEstBlock(templates[120], 5.0, 1, 0.9), oracle.CorrEstBlock(templates[120], 5.0, 1, 0.9)
    x = stream(137 * 6 + 2048, 8, burst_at=[100, 500], tmpl=templates[120])
    pos = 0
    for n, swap in ((137, None), (373, templates[140]), (274, templates[120]), (137, None)):
        if swap is not None:
            ref.set_symbols(swap)
            ora.set_symbols(swap)
        L = len(ref.symbols())
        a = ref.work(n, x[pos:pos + n + L], nitems_written=pos, two_ports=True)
        b = ora.work(n, x[pos:pos + n + L], nitems_written=pos, two_ports=True)
        assert np.array_equal(a[1], b[1])
        same_tags(a[3], b[3])
        pos += n


# ------------------------------------------------------ msk_timing_recovery_cc

def gmsk(nbits, seed):
    from gr_ais_b200 import synth
    rng = np.random.default_rng(seed)
    x = synth.gmsk_modulate(rng.integers(0, 2, nbits)).astype(np.complex64)
    return x + (0.05 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)


def test_msk_ctor_setters_forecast_and_exceptions(oracle):
    ref, ora = R.MskBlock(5.0, 0.04, 0.01, 1), oracle.MskBlock(5.0, 0.04, 0.01, 1)
    assert ref.get_sps() == 2.5 == ora.state.sps and ref.get_gain() == np.float32(0.04)
    assert ref.get_limit() == np.float32(0.01) and ref.relative_rate() == pytest.approx(0.2)
    for nout in (0, 1, 100, 4096, 24576):
        assert ref.forecast(nout) == ora.forecast(nout)
    for bad in ((5.0, 0.0, 0.01, 1), (5.0, -1.0, 0.01, 1), (5.0, 0.04, 0.01, 3), (5.0, 0.04, 0.01, 0)):
        with pytest.raises(IndexError):
            R.MskBlock(*bad)
        with pytest.raises(IndexError):
            oracle.MskBlock(*bad)
    with pytest.raises(IndexError):
        ref.set_gain(0.0)
    ref.set_sps(5.208)
    assert ref.get_sps() == np.float32(np.float32(5.208) / 2.0)


@pytest.mark.parametrize("osps", [1, 2])
def test_msk_general_work_quirks_streamed_in_ragged_calls(oracle, osps):
    """tag reset (:141-158) incl. negative centre => iidx-- (:148-153), NaN tag dropped without a
    reset (:144-147), only tags[0] examined (a stale tag blocks the ones behind it), off-key
    tags ignored, unclipped err on even steps, absolute limit, 3*d_sps items left unconsumed."""
    x = gmsk(800, 4)
    tags = mk_tags([(333, TAG["time_est"], 0.31), (334, TAG["phase_est"], 0.5),
                    (900, TAG["time_est"], -0.42), (1500, TAG["time_est"], np.nan),
                    (1501, TAG["time_est"], 0.07), (2500, TAG["corr_start"], 9.0),
                    (2600, TAG["time_est"], 0.999), (2601, TAG["time_est"], 0.5),
                    (2602, TAG["time_est"], -0.999), (3100, TAG["time_est"], -1e-9)])
    ref, ora = R.MskBlock(5.0, 0.04, 0.01, osps), oracle.MskBlock(5.0, 0.04, 0.01, osps)
    pos = 0
    for avail, nout in ((800, 1000), (1700, 60), (1702, 1000), (2400, 1000), (2400, 1000), (len(x), 4000)):
        chunk = x[pos:avail]
        a = ref.general_work(nout, chunk, tags, nitems_read=pos)
        b = ora.general_work(nout, chunk, tags, nitems_read=pos)
        assert a[3] == b[3] and len(a[0]) == len(b[0])
        for u, v in zip(a[:3], b[:3]):
            assert np.array_equal(u, v)
        pos += a[3]
    assert pos > len(x) - 20


def test_msk_negative_centre_on_the_first_item_reads_in_minus_one(oracle):
    """offset == nitems_read with a negative centre makes iidx = -1 (:148-153): the block reads the
    item before its read pointer -- the last item consumed by the previous call."""
    x = gmsk(300, 9)
    ref, ora = R.MskBlock(5.0, 0.04, 0.01, 1), oracle.MskBlock(5.0, 0.04, 0.01, 1)
    a = ref.general_work(1000, x[:700], None, 0)
    b = ora.general_work(1000, x[:700], None, 0)
    assert a[3] == b[3]
    pos = a[3]
    tags = mk_tags([(pos, TAG["time_est"], -0.3)])
    a = ref.general_work(1000, x[pos:], tags, pos)
    b = ora.general_work(1000, x[pos:], tags, pos)
    assert a[3] == b[3]
    for u, v in zip(a[:3], b[:3]):
        assert np.array_equal(u, v)


def test_msk_gain_and_limit_changes_midstream(oracle):
    x = gmsk(600, 12)
    ref, ora = R.MskBlock(5.0, 0.25, 0.3, 1), oracle.MskBlock(5.0, 0.25, 0.3, 1)
    a = ref.general_work(5000, x[:1500], None, 0)
    b = ora.general_work(5000, x[:1500], None, 0)
    assert a[3] == b[3] and np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.ptp(a[2]) > 0.2            # the loop really moves mu with this gain


# ------------------------------------------------------------------- freqest

@pytest.mark.parametrize("fs,dr,fftlen", [(48000.0, 9600, 1024), (48000.0, 9600, 256),
                                          (50000.0, 9600, 1024), (250000.0 / 5, 9600, 4096)])
def test_freqest_work_with_maxpos_carry_over(oracle, fs, dr, fftlen):
    """maxpos is only reset per work() call, not per vector (freqest_impl.cc:67-68,74): an all-zero
    vector repeats the previous vector's estimate, and gives bin 0 when it comes first."""
    rng = np.random.default_rng(3)
    spec = (rng.standard_normal((7, fftlen)) + 1j * rng.standard_normal((7, fftlen))).astype(np.complex64)
    spec[2] = 0
    spec[5, 300 % fftlen] = 80.0
    a = R.freqest_work(spec, fs, dr, fftlen)
    b, _ = oracle.freqest_work(spec, fs, dr, fftlen)
    assert np.array_equal(a, b) and a[2] == a[1]
    z = np.zeros((2, fftlen), np.complex64)
    assert np.array_equal(R.freqest_work(z, fs, dr, fftlen), oracle.freqest_work(z, fs, dr, fftlen)[0])


def test_freqest_std_abs_is_the_canonical_hypot(oracle):
    """std::abs(gr_complex) in the reference is this libm's hypotf; the oracle (and the CUDA kernel)
    evaluate sqrt(re^2 + im^2) in double -- same bits over a wide dynamic range."""
    rng = np.random.default_rng(5)
    spec = ((rng.standard_normal((40, 1024)) + 1j * rng.standard_normal((40, 1024))) *
            10.0 ** rng.uniform(-12, 12, (40, 1))).astype(np.complex64)
    assert np.array_equal(R.freqest_work(spec), oracle.freqest_work(spec)[0])


# -------------------------------------------------------------------- invert

def test_invert(oracle):
    b = np.arange(256, dtype=np.uint8)
    assert np.array_equal(R.invert(b), oracle.invert(b))
    assert np.array_equal(R.invert(b), (b ^ 1) & 1)


# --------------------------------------------------------------- pdu_to_nmea

def _dearmour(payload, npad):
    """6-bit ASCII armour -> bytes (ITU-R M.1371 / NMEA 0183 AIVDM)"""
    bits = []
    for ch in payload:
        v = ord(ch) - 48
        if v > 40:
            v -= 8
        bits += [(v >> (5 - k)) & 1 for k in range(6)]
    if npad:
        bits = bits[:-npad]
    return np.packbits(np.array(bits, np.uint8)).tobytes()


def test_pdu_to_nmea_reproduces_public_sentences(oracle):
    """the reference's formatter itself regenerates public AIVDM sentences from their payloads"""
    with open(os.path.join(HERE, "golden", "aivdm_kat.json")) as f:
        kat = json.load(f)
    for s in kat["single"]:
        f = s[:s.index("*")].split(",")
        pdu = _dearmour(f[5], int(f[6]))
        assert R.pdu_to_nmea(pdu, f[4]) == s == oracle.pdu_to_nmea(pdu, f[4])
    a, b = kat["multi"][0]
    fa, fb = a[:a.index("*")].split(","), b[:b.index("*")].split(",")
    pdu = _dearmour(fa[5] + fb[5], int(fb[6]))
    got = R.pdu_to_nmea(pdu, fa[4])
    assert got == oracle.pdu_to_nmea(pdu, fa[4]) and got.count("\n") == 1
    assert [g.split(",")[5] for g in got.split("\n")] == [fa[5], fb[5]]


def test_pdu_to_nmea_random_lengths_designators_and_fragmentation(oracle):
    rng = np.random.default_rng(8)
    for n in list(range(1, 64)) + [84, 85, 126, 127, 168, 248]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for des in ("A", "B", "AB"):
            a, b = R.pdu_to_nmea(data, des), oracle.pdu_to_nmea(data, des)
            assert a == b, (n, des)
    # the padding quirk (:75-77): the last group is shifted again inside a uint8_t
    assert R.pdu_to_nmea(b"\xff", "A") == oracle.pdu_to_nmea(b"\xff", "A")
    assert R.pdu_to_nmea(b"\xff\xff", "A").split(",")[5][2] == chr((240 - 256 + 48) & 0xFF)


# ------------------------------------------------------------ the whole chain

def load(name):
    return np.load(os.path.join(HERE, "golden", name))


def test_chain_on_the_reference_blocks_equals_golden_vectors_and_oracle(oracle):
    z = load("chain_kat.npz")
    for name in ("l120", "l140"):
        x, tmpl = z[name + "_iq"], z[name + "_template"]
        a = oracle.demod_chain(x, tmpl, debug=True, blocks=R.blocks())
        b = oracle.demod_chain(x, tmpl, debug=True)
        assert np.array_equal(a["bits"], z[name + "_bits"])
        same_tags(a["tags"], z[name + "_tags"])
        for k in ("fhat", "mixed", "agc", "corr", "mag", "sym", "err", "mu", "soft", "bits"):
            assert np.array_equal(a[k], b[k]), k
        assert (a["n1"], a["n2"], a["consumed"]) == (b["n1"], b["n2"], b["consumed"])


@pytest.mark.parametrize("kind", ["north_star", "intended", "reference"])
def test_chain_on_fresh_synthetic_records(oracle, kind):
    from gr_ais_b200 import synth
    from gr_ais_b200.ais_demod import preamble_template
    tmpl = preamble_template(kind)
    for seed, snr in ((31, 20.0), (32, 8.0), (33, 3.0)):
        x, _ = synth.make_record(seed, n=24000, nbursts=3, snr_db=snr, random_impairments=True)
        a = oracle.demod_chain(x, tmpl, blocks=R.blocks())
        b = oracle.demod_chain(x, tmpl)
        assert np.array_equal(a["bits"], b["bits"])
        same_tags(a["tags"], b["tags"])
        assert len(a["bits"]) > 4000


def test_stream_on_the_reference_blocks_in_ragged_pieces(oracle):
    z = load("chain_kat.npz")
    x, tmpl = z["l120_iq"], z["l120_template"]
    a, b = oracle.DemodStream(tmpl, blocks=R.blocks()), oracle.DemodStream(tmpl)
    cuts = [0, 1, 1, 700, 5000, 5000, 9999, 16000, len(x)]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        (ab, at), (bb, bt) = a.work(x[lo:hi]), b.work(x[lo:hi])
        assert np.array_equal(ab, bb)
        same_tags(at, bt)


def test_stream_set_symbols_hot_swap(oracle, templates):
    z = load("chain_kat.npz")
    x = z["l120_iq"]
    a, b = oracle.DemodStream(templates[120], blocks=R.blocks()), oracle.DemodStream(templates[120])
    new = np.conj(templates[120])[::-1].copy() * np.complex64(0.5)
    for k, (lo, hi) in enumerate(((0, 6000), (6000, 11000), (11000, len(x)))):
        if k == 1:
            a.set_symbols(new)
            b.set_symbols(new)
        (ab, at), (bb, bt) = a.work(x[lo:hi]), b.work(x[lo:hi])
        assert np.array_equal(ab, bb)
        same_tags(at, bt)


def test_batch_on_the_reference_blocks(oracle):
    from gr_ais_b200 import synth
    from gr_ais_b200.ais_demod import preamble_template
    tmpl = preamble_template("north_star")
    x = np.stack([synth.make_record(50 + c, n=12288, nbursts=2, snr_db=15.0)[0] for c in range(6)])
    a = oracle.demod_chain_batch(x, tmpl, nthreads=2, blocks=R.blocks())
    b = oracle.demod_chain_batch(x, tmpl, nthreads=2)
    for u, v in zip(a, b):
        if u.dtype.names:
            for c in range(len(x)):
                same_tags(u[c, :a[3][c]], v[c, :b[3][c]])
        else:
            assert np.array_equal(u, v)
