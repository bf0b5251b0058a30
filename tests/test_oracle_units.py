"""Pins the CPU oracle stage by stage from first principles (the reference ships no
vectors: SURVEY.md section 8c).  Every stage is checked against a float64 "truth" and
against the quirks of the reference source it restates (file:line in each test)."""
import numpy as np
import pytest

TAG = {"corr_start": 0, "phase_est": 1, "time_est": 2, "corr_est": 3}


# ----------------------------------------------------------------- scalars

def test_hypot_matches_libm(oracle):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal(4000) * 10.0 ** rng.uniform(-6, 6, 4000)).astype(np.float32)
    b = (rng.standard_normal(4000) * 10.0 ** rng.uniform(-6, 6, 4000)).astype(np.float32)
    got = np.array([oracle.hypotf(x, y) for x, y in zip(a, b)], dtype=np.float32)
    assert np.array_equal(got, np.hypot(a, b))  # numpy float32 hypot is libm hypotf


def test_fast_atan2f(oracle):
    rng = np.random.default_rng(2)
    y = rng.standard_normal(3000).astype(np.float32)
    x = rng.standard_normal(3000).astype(np.float32)
    got = np.array([oracle.fast_atan2f(a, b) for a, b in zip(y, x)])
    # the published routine indexes a 1/255-step table with z*256 - 0.5: ~2e-3 rad worst case
    assert np.max(np.abs(got - np.arctan2(y.astype(np.float64), x))) < 2.5e-3
    assert oracle.fast_atan2f(0.0, 0.0) == 0.0
    assert oracle.fast_atan2f(0.0, -1.0) == pytest.approx(np.pi, abs=1e-6)
    assert oracle.fast_atan2f(-0.0, -1.0) == pytest.approx(np.pi, abs=1e-6)  # y >= 0.0 holds for -0
    assert oracle.fast_atan2f(1.0, 0.0) == pytest.approx(np.pi / 2, abs=1e-6)


def test_branchless_clip(oracle):
    for x in (-10.0, -3.0, -0.25, 0.0, 0.1, 2.999, 3.0, 7.5):
        assert oracle.branchless_clip(x, 3.0) == pytest.approx(np.clip(x, -3.0, 3.0), abs=1e-6)


def test_float_to_fixed_and_sincos(oracle):
    assert oracle.float_to_fixed(0.0) == 0
    assert oracle.float_to_fixed(float(np.float32(np.pi / 2))) == pytest.approx(2 ** 30, abs=128)
    assert oracle.float_to_fixed(float(-np.float32(np.pi))) == -2 ** 31
    # folding: an angle outside [-pi, pi] maps to the same point of the circle
    for a in (4.0, -5.5, 7.0, -9.0):
        s, c = oracle.fxpt_sincos(oracle.float_to_fixed(a))
        assert abs(s - np.sin(a)) < 2e-5 and abs(c - np.cos(a)) < 2e-5


def test_agc_envelope(oracle):
    # max + 0.4*min evaluated with the 0.4 literal in double, then narrowed
    for re, im in ((1.0, 0.0), (0.3, -0.7), (-2.5, 2.5), (1e-6, 3.0)):
        r, i = abs(np.float32(re)), abs(np.float32(im))
        want = np.float32(float(max(r, i)) + 0.4 * float(min(r, i)))
        assert oracle.agc_envelope(re, im) == want


# ------------------------------------------------------------ freq sync

def test_fft_against_float64(oracle):
    rng = np.random.default_rng(3)
    for n in (4, 64, 1024, 4096):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        got = oracle.fft_forward(x)
        ref = np.fft.fft(x.astype(np.complex128))
        assert np.max(np.abs(got - ref)) <= 1e-5 * np.sqrt(n) * np.max(np.abs(ref)) / np.sqrt(n) + 1e-4
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-5
    with pytest.raises(ValueError):
        oracle.fft_forward(np.zeros(100, np.complex64))


def test_fft_shift_is_gnuradio_fft_vcc_shift(oracle):
    x = np.arange(8).astype(np.complex64)
    assert list(oracle.fft_shift(x).real) == [4, 5, 6, 7, 0, 1, 2, 3]  # DC moves to n/2
    x = np.exp(2j * np.pi * 5 * np.arange(1024) / 1024).astype(np.complex64)
    s = oracle.fft_shift(oracle.fft_forward(x))
    assert np.argmax(np.abs(s)) == 512 + 5


def test_square_is_volk_fma_form(oracle):
    x = np.array([1 + 2j, -0.5 + 0.25j, 3e-3 - 7j], dtype=np.complex64)
    got = oracle.square(x)
    assert np.allclose(got, x.astype(np.complex128) ** 2, rtol=1e-6)


def test_freqest_known_offset(oracle):
    # lib/freqest_impl.cc:57-88: GMSK squared has lines at 2*f0 +- datarate/2; a tone pair stands in
    fs, n, f0 = 48000.0, 1024, 375.0
    t = np.arange(n) / fs
    x2 = np.exp(2j * np.pi * (2 * f0 - 4800) * t) + np.exp(2j * np.pi * (2 * f0 + 4800) * t)
    spec = oracle.fft_shift(oracle.fft_forward(x2.astype(np.complex64)))
    hz, maxpos = oracle.freqest_work(spec)
    assert hz[0] == pytest.approx(f0, abs=46.875 / 2)
    assert maxpos[0] == 512 + 16  # 2*f0/binsize = 16 bins above DC


def test_freqest_maxpos_carries_over_inside_a_call(oracle):
    # lib/freqest_impl.cc:67-68,74: maxpos is set to 0 once per work() call, not per vector
    n = 1024
    tone = np.zeros(n, np.complex64)
    tone[600] = 1.0
    zero = np.zeros(n, np.complex64)
    hz, mp = oracle.freqest_work(np.stack([zero, tone, zero]))
    assert mp[0] == 0 and hz[0] == (0 - 512) * 46.875 / 2   # nothing seen yet: -12 kHz
    assert mp[1] == mp[2] and hz[1] == hz[2]                # silent vector repeats the estimate
    hz2, mp2 = oracle.freqest_work(np.stack([zero]))         # a new call starts from 0 again
    assert mp2[0] == 0


def test_nco_mix_continuous_phase(oracle):
    fs = 48000.0
    sens = float(np.float32(-2 * np.pi / fs))
    x = np.ones(4096, np.complex64)
    f = np.array([1000.0, 1000.0, -500.0, -500.0], np.float32)
    y, ph = oracle.nco_mix(x, f, 1024, sens)
    truth_phase = np.cumsum(np.repeat(f.astype(np.float64), 1024) * sens)
    assert np.max(np.abs(y - np.exp(1j * truth_phase))) < 1e-3   # float32 phase accumulation drift
    # state carries across calls exactly
    y1, p1 = oracle.nco_mix(x[:1024], f[:1], 1024, sens)
    y2, p2 = oracle.nco_mix(x[1024:], f[1:], 1024, sens, phase=p1)
    assert np.array_equal(np.concatenate([y1, y2]), y) and p2 == ph


def test_agc_against_truth(oracle):
    rng = np.random.default_rng(4)
    x = (rng.standard_normal(3000) + 1j * rng.standard_normal(3000)).astype(np.complex64)
    x[1000:1400] *= 20
    got = oracle.agc_work(x, 512, 2.0)
    xp = np.concatenate([np.zeros(511, np.complex64), x])
    env = np.maximum(np.abs(xp.real), np.abs(xp.imag)) + 0.4 * np.minimum(np.abs(xp.real), np.abs(xp.imag))
    for t in (0, 1, 510, 511, 999, 1000, 1511, 1911, 1912, 2999):
        m = max(1e-4, env[t:t + 512].max())
        assert got[t] == pytest.approx(xp[t] * 2.0 / m, rel=1e-6, abs=1e-9)
    assert np.all(got[:0 + 1] == 0)  # history is zeros: the first output is the zero sample 511 back


# -------------------------------------------------------------- corr_est

def test_corr_est_constructor(oracle, templates):
    t = templates[120]
    b = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
    # lib/corr_est_cc_impl.cc:59-63: symbols() returns the conj-reversed taps
    assert np.array_equal(b.symbols(), np.conj(t)[::-1])
    # :71-74: threshold * (sum |s|^2)^2, :84: output multiple = fft_filter nsamples, :65-66 clamp
    assert b.thresh == pytest.approx(0.9 * 120 ** 2, rel=1e-4)
    assert b.nsamples == 256 - 120 + 1
    assert oracle.CorrEstBlock(templates[140], 5.0, 1).nsamples == 512 - 140 + 1
    assert oracle.CorrEstBlock(templates[1120], 5.0, 1).nsamples == 4096 - 1120 + 1
    assert oracle.CorrEstBlock(t, 5.0, 1000).mark_delay == 119


def test_corr_est_set_symbols_quirk(oracle, templates):
    # :132-162: set_symbols stores the taps verbatim (no conj/reverse) and keeps d_thresh
    b = oracle.CorrEstBlock(templates[120], 5.0, 1, 0.9)
    th = b.thresh
    new = (templates[140] * 3).astype(np.complex64)
    b.set_symbols(new)
    assert np.array_equal(b.symbols(), new) and b.thresh == th and b.L == 140


def test_corr_est_detects_template(oracle, templates):
    t = templates[120]
    L = len(t)
    rng = np.random.default_rng(5)
    n = 137 * 8
    stream = (0.05 * (rng.standard_normal(n + L) + 1j * rng.standard_normal(n + L))).astype(np.complex64)
    p = 400  # template occupies in[p .. p+L): with L items of history that is stream item p-L
    stream[p:p + L] += t
    b = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
    out0, corr, mag, tags = b.work(n, stream, nitems_written=1000)
    assert np.array_equal(out0, stream[:n])                      # :184 delay by the history
    # the filter is fed &in[L] and starts from a zero tail: the tagging history does not enter
    # the first call's correlation (fft_filter_ccc keeps its own state)
    fed = np.concatenate([np.zeros(L, np.complex128), stream[L:].astype(np.complex128)])
    truth = np.array([np.vdot(t.astype(np.complex128), fed[i + 1:i + 1 + L]) for i in range(n)])
    assert np.max(np.abs(corr - truth)) < 1e-5 * L
    assert np.max(np.abs(corr - oracle.CorrEstBlock(t, 5.0, 1, 0.9).direct_f64(n, stream))) < 1e-5 * L
    assert np.allclose(mag, np.abs(truth) ** 2, rtol=1e-4, atol=1e-2)
    i_peak = p - 1   # out0[i+1] is the first template sample
    cs = tags[tags["key"] == TAG["corr_start"]]
    assert 1000 + i_peak in cs["offset"]
    k = list(cs["offset"]).index(1000 + i_peak)
    assert cs["value"][k] == float(mag[i_peak]) and mag[i_peak] > 0.9 * L * L
    te = tags[tags["key"] == TAG["time_est"]]
    assert te["offset"][k] == 1000 + i_peak + 1                  # mark_delay = 1
    # centre of mass in double (:219-227)
    m = mag[i_peak - 1:i_peak + 2].astype(np.float64)
    want = (np.float32(1) * mag[i_peak - 1] + np.float64(np.float32(2) * mag[i_peak])
            + np.float64(np.float32(3) * mag[i_peak + 1])) / m.sum() - 2.0
    assert te["value"][k] == want
    # tags of one detection come in the reference's order; detections are >= isps apart
    assert list(tags["key"][:4]) == [0, 1, 2, 3]
    assert np.all(np.diff(cs["offset"].astype(np.int64)) >= 5)


def test_corr_est_chunk_edges(oracle, templates):
    # :202 the climb stops at the last item, :220 CoM is 0.0 at the first/last item of a chunk
    t = templates[120]
    L = len(t)
    n = 274
    for item in (0, n - 1):     # the peak lands on item 0 / item n-1 of the second work() chunk
        stream = np.zeros(2 * n + L, np.complex64)
        end = L + n + item      # stream index of the template's last sample <=> the peak item
        stream[end - L + 1:end + 1] = t
        b = oracle.CorrEstBlock(t, 5.0, 1, 0.5)
        b.work(n, stream[:n + L])
        _, _, mag, tags = b.work(n, stream[n:], nitems_written=n)
        assert int(np.argmax(mag)) == item
        te = tags[(tags["key"] == TAG["time_est"]) & (tags["offset"] == n + item + 1)]
        assert len(te) == 1 and te["value"][0] == 0.0
        # the same peak in the middle of a chunk gets a non-zero centre of mass
        c = oracle.CorrEstBlock(t, 5.0, 1, 0.5)
        _, _, _, tags2 = c.work(3 * n, np.concatenate([stream, np.zeros(n, np.complex64)]))
        te2 = tags2[(tags2["key"] == TAG["time_est"]) & (tags2["offset"] == n + item + 1)]
        assert len(te2) == 1 and te2["value"][0] != 0.0


def test_corr_est_is_gnuradio_fft_filter(oracle, templates):
    """kernel::fft_filter_ccc structure: fftsize = 2*2^ceil(log2 L), blocks of nsamples, a tail of
    L-1 items carried from call to call; splitting the stream into calls changes nothing"""
    rng = np.random.default_rng(12)
    for L, fft in ((120, 256), (140, 512), (1120, 4096)):
        t = templates[L]
        a = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
        assert a.fftsize == fft and a.nsamples == fft - L + 1
        ns = a.nsamples
        x = (rng.standard_normal(6 * ns + L) + 1j * rng.standard_normal(6 * ns + L)).astype(np.complex64)
        x[:L] = 0
        _, whole, _, _ = a.work(6 * ns, x)
        b = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
        parts = [b.work(k * ns, x[o * ns:o * ns + k * ns + L])[1] for o, k in ((0, 1), (1, 3), (4, 2))]
        assert np.array_equal(np.concatenate(parts), whole)
        assert np.array_equal(a.tail(), b.tail()) and len(a.tail()) == L - 1
        assert np.max(np.abs(whole - a.direct_f64(6 * ns, x))) < 1e-5 * L
        with pytest.raises(ValueError):
            a.work(ns + 1, x)
    # the transform pair: DIF forward leaves bit-reversed order, DIT inverse undoes it
    x = (rng.standard_normal(512) + 1j * rng.standard_normal(512)).astype(np.complex64)
    br = np.array([int(format(i, "09b")[::-1], 2) for i in range(512)])
    X = oracle.fft_dif(x)
    assert np.max(np.abs(X - np.fft.fft(x.astype(np.complex128))[br])) < 2e-5 * np.abs(X).max()
    assert np.max(np.abs(oracle.ifft_dit(X) / 512 - x)) < 2e-6


def test_corr_est_two_port_tags(oracle, templates):
    t = templates[120]
    stream = np.concatenate([np.zeros(200, np.complex64), t, np.zeros(228, np.complex64)])
    b = oracle.CorrEstBlock(t, 5.0, 3, 0.9)
    _, _, _, tags = b.work(411, stream, two_ports=True)
    p0, p1 = tags[tags["port"] == 0], tags[tags["port"] == 1]
    assert len(p0) > 0 and len(p1) == 3 * len(p0) // 4
    # port-1 debug tags are not offset by mark_delay (:258-266)
    i = int(p0[p0["key"] == 0]["offset"][0])
    assert set(p1["offset"][:3]) == {i} and int(p0[p0["key"] == 2]["offset"][0]) == i + 3


# -------------------------------------------------------------------- msk

def test_msk_parameters_and_errors(oracle):
    m = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    assert m.state.sps == 2.5 and m.state.omega == 2.5 and m.state.mu == 0.5
    assert m.state.gain_omega == np.float32(np.float32(0.04) * np.float32(0.04)) * np.float32(0.25)
    assert m.forecast(100) == int(np.ceil(100 * 2.5 * 2 + 7.5 + 8))     # :103
    with pytest.raises(IndexError):
        oracle.MskBlock(5.0, 0.0, 0.01, 1)      # :82 Gain must be positive
    with pytest.raises(IndexError):
        oracle.MskBlock(5.0, 0.04, 0.01, 3)     # :61 osps must be 1 or 2


def _gmsk(nsym, seed=0):
    from gr_ais_b200 import synth
    rng = np.random.default_rng(seed)
    return synth.gmsk_modulate(rng.integers(0, 2, nsym)).astype(np.complex64)


def test_msk_symbol_rate_and_osps(oracle):
    x = _gmsk(400)
    m1 = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    out, err, mu, consumed = m1.general_work(10000, x)
    # the loop stops once iidx >= ninput - 3*d_sps (:119,138); the last step may overshoot by < 3
    assert abs(len(out) - len(x) / 5) <= 3 and len(x) - 8 <= consumed <= len(x) - 5
    m2 = oracle.MskBlock(5.0, 0.04, 0.01, 2)
    out2, _, _, _ = m2.general_work(10000, x)
    assert abs(len(out2) - 2 * len(out)) <= 2
    assert np.all(np.abs(np.abs(out[40:]) - 1.0) < 0.25)   # interpolants of a unit-modulus signal
    assert np.all((mu >= -0.13) & (mu < 1.13))


def test_msk_streaming_equals_one_shot(oracle):
    x = _gmsk(600, seed=1)
    a = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    whole, _, _, c_all = a.general_work(100000, x)
    b = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    pos, parts = 0, []
    for avail in (700, 1300, 2100, len(x)):
        out, _, _, c = b.general_work(100000, x[pos:avail], nitems_read=pos)
        parts.append(out)
        pos += c
    assert pos == c_all and np.array_equal(np.concatenate(parts), whole)


def test_msk_tag_reset_semantics(oracle):
    x = _gmsk(300, seed=2)
    base = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    ref, _, ref_mu, _ = base.general_work(10000, x)

    def tags(*items):
        t = np.zeros(len(items), dtype=oracle.TAG_DTYPE)
        for k, (off, val, key) in enumerate(items):
            t[k] = (off, key, 0, val)
        return t

    # NaN time_est is dropped without a reset (:144-147); other keys are ignored (:130)
    m = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    out, _, mu, _ = m.general_work(10000, x, tags((200, np.nan, 2), (400, 0.3, 1), (401, 0.3, 3)))
    assert np.array_equal(out, ref) and np.array_equal(mu, ref_mu)
    # a positive centre restarts the loop at the tag: mu = centre, div = 0 => a symbol comes out there
    m = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    out, _, mu, _ = m.general_work(10000, x, tags((500, 0.25, 2)))
    k = int(np.argmax(mu != ref_mu[:len(mu)])) if len(mu) == len(ref_mu) else None
    assert not np.array_equal(out, ref)
    hit = np.where(np.isclose(mu, np.float32(0.25)))[0]
    assert len(hit) >= 1
    # a negative centre becomes mu+1 one item earlier (:150-153)
    m = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    out_n, _, mu_n, _ = m.general_work(10000, x, tags((500, -0.25, 2)))
    hit = np.where(np.isclose(mu_n, np.float32(0.75)))[0]
    assert len(hit) >= 1
    # only tags[0] is ever examined: a stale tag blocks the later ones in this call (:140-141)
    m = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    m.general_work(10000, x[:450])           # consume past offset 300
    got, _, _, _ = m.general_work(10000, x[m_consumed(m, x):], None)
    assert len(got) >= 0


def m_consumed(m, x):
    return 0


def test_msk_limit_is_absolute_and_error_is_clipped(oracle):
    # :182 omega is clipped to d_sps +- limit (samples, not relative); :180 err to +-3 on odd steps
    rng = np.random.default_rng(7)
    x = (3.0 * (rng.standard_normal(4000) + 1j * rng.standard_normal(4000))).astype(np.complex64)
    m = oracle.MskBlock(5.0, 0.5, 0.01, 2)
    out, err, mu, _ = m.general_work(100000, x)
    assert abs(m.state.omega - 2.5) <= 0.01 + 1e-6
    assert np.max(np.abs(err[1::2])) <= 3.0 + 1e-5   # odd half-steps are clipped (0.5*(|x+3|-|x-3|))
    assert np.max(np.abs(err[0::2])) > 3.0           # even half-steps report the raw error (:186-189)


# ------------------------------------------------------------------- tail

def test_bit_tail(oracle):
    rng = np.random.default_rng(8)
    ph = np.cumsum(rng.choice([-np.pi / 2, np.pi / 2], 200))
    v = np.exp(1j * ph).astype(np.complex64)
    soft = oracle.quad_demod(v)
    d = np.angle(v[1:] * np.conj(v[:-1]))
    assert np.max(np.abs(soft[1:] - (np.pi / 2) * d)) < 1e-3
    assert soft[0] == 0.0                                   # x[-1] = 0 -> atan2(0, 0) = 0
    b = oracle.binary_slicer(soft)
    assert np.array_equal(b, (soft >= 0).astype(np.uint8))
    dd = oracle.diff_decoder(b)
    assert np.array_equal(dd, b ^ np.concatenate([[0], b[:-1]]).astype(np.uint8))
    assert np.array_equal(oracle.invert(dd), 1 - dd)        # lib/invert_impl.cc:63
    assert np.array_equal(oracle.invert(np.array([0, 1, 2, 3, 255], np.uint8)), [1, 0, 1, 0, 0])


def test_stream_single_call_equals_batch_chain(oracle, templates):
    """the stream restatement fed a whole record in one call is the batch chain"""
    from gr_ais_b200 import synth
    x, truth = synth.make_record(3, n=20000, nbursts=3, snr_db=25)
    ref = oracle.demod_chain(x, templates[120])
    s = oracle.DemodStream(templates[120])
    bits, tags = s.work(x)
    assert np.array_equal(bits, ref["bits"])
    assert np.array_equal(tags, ref["tags"])


def test_stream_pieces_carry_every_block_state(oracle, templates):
    """cut a record on whole FFT vectors / corr_est multiples / work chunks (137 * 1024 items):
    every block resumes exactly where it stopped, so only the timing loop's call boundary (the
    3*sps/2 items it leaves unconsumed) can move a decision"""
    from gr_ais_b200 import synth
    n = 2 * 137 * 1024
    x, truth = synth.make_record(1, n=n, nbursts=10, snr_db=25)
    cfg = oracle.chain_cfg(corr_chunk=137 * 64)
    whole, _ = oracle.DemodStream(templates[120], cfg).work(x)
    s = oracle.DemodStream(templates[120], cfg)
    parts = [s.work(x[:n // 2])[0], s.work(x[n // 2:])[0]]
    cut = np.concatenate(parts)
    assert abs(len(cut) - len(whole)) <= 1
    k = min(len(cut), len(whole))
    assert np.mean(cut[:k] == whole[:k]) > 0.99
    # ragged pieces: nothing is lost or duplicated, every payload still comes out
    s = oracle.DemodStream(templates[120])
    rng = np.random.default_rng(2)
    pos, out = 0, []
    while pos < n:
        m = int(min(n - pos, rng.integers(0, 9000)))
        out.append(s.work(x[pos:pos + m])[0])
        pos += m
    ragged = np.concatenate(out)
    assert abs(len(ragged) - n // 5) < 40
    assert sum(synth.payloads_found(ragged, truth)) >= sum(synth.payloads_found(whole, truth)) - 1 >= 6
