"""Seeded random configurations: the oracle's restatements against the reference's own classes
(oracle/_ref, see tests/test_ref_pin.py) far from the directed cases -- rates, gains, limits,
thresholds, template lengths, tag lists (stale, off-key, NaN, negative centres) and ragged call
sizes drawn at random.  Everything must be identical bit for bit."""
import numpy as np
import pytest

from oracle import ref as R
from test_ref_pin import TAG, gmsk, mk_tags, same_tags

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and no reference tree")


@pytest.mark.parametrize("seed", range(24))
def test_msk_random_rates_tags_and_call_sizes(oracle, seed):
    rng = np.random.default_rng(9000 + seed)
    # sps >= 16/3 keeps the interpolator's 8 items inside ninput_items[0]: the loop runs while
    # iidx < ninput - 3 sps/2 and reads in[iidx .. iidx + 7], so at sps 4 the last step of a call
    # reads one item past what the scheduler promised (whatever the buffer holds there)
    sps = float(rng.choice([5.0, 5.0, 5.208, 5.5, 6.0, 8.0]))
    gain = float(rng.choice([0.01, 0.04, 0.1, 0.175]))
    limit = float(rng.choice([0.005, 0.01, 0.05, 0.2]))
    osps = int(rng.integers(1, 3))
    x = gmsk(500, 100 + seed)
    if sps != 5.0:  # resample by linear interpolation: any smooth complex stream will do
        t = np.arange(0, len(x) - 1, 5.0 / sps)
        i = t.astype(int)
        f = (t - i).astype(np.float32)
        x = ((1 - f) * x[i] + f * x[i + 1]).astype(np.complex64)
    n = len(x)
    offs = np.sort(rng.integers(0, n, int(rng.integers(0, 14))))
    items = []
    for o in offs:
        key = TAG["time_est"] if rng.random() < 0.75 else int(rng.integers(0, 4))
        val = float(rng.uniform(-0.999, 0.999)) if rng.random() < 0.9 else float("nan")
        items.append((int(o), key, val))
    tags = mk_tags(items)
    ref, ora = R.MskBlock(sps, gain, limit, osps), oracle.MskBlock(sps, gain, limit, osps)
    pos = 0
    for _ in range(40):
        avail = min(n, pos + int(rng.integers(0, 900)))
        nout = int(rng.integers(1, 400))
        chunk = x[pos:avail]
        try:
            a = ref.general_work(nout, chunk, tags, nitems_read=pos)
        except RuntimeError:
            with pytest.raises(RuntimeError):
                ora.general_work(nout, chunk, tags, nitems_read=pos)
            return
        b = ora.general_work(nout, chunk, tags, nitems_read=pos)
        assert a[3] == b[3] and len(a[0]) == len(b[0]), (seed, pos)
        for u, v in zip(a[:3], b[:3]):
            assert np.array_equal(u.view(np.uint32), v.view(np.uint32)), (seed, pos)
        pos += a[3]
        if pos >= n:
            break


@pytest.mark.parametrize("seed", range(16))
def test_corr_est_random_templates_thresholds_and_chunks(oracle, seed):
    rng = np.random.default_rng(7000 + seed)
    L = int(rng.choice([5, 9, 16, 31, 64, 100, 120, 140, 257]))
    tmpl = np.exp(2j * np.pi * rng.uniform(0, 1, L)).astype(np.complex64)
    thr = float(rng.choice([0.2, 0.5, 0.8, 0.9]))
    md = int(rng.integers(0, 2 * L))
    ref, ora = R.CorrEstBlock(tmpl, 5.0, md, thr), oracle.CorrEstBlock(tmpl, 5.0, md, thr)
    ns = ora.nsamples
    assert ref.nsamples == ns and ref.hints()["output_multiple"] == ns
    blocks = [int(rng.integers(1, 5)) for _ in range(4)]
    total = sum(blocks) * ns
    x = (0.2 * (rng.standard_normal(total + L) + 1j * rng.standard_normal(total + L))).astype(np.complex64)
    for _ in range(int(rng.integers(1, 6))):
        p = int(rng.integers(0, total))
        seg = x[p:p + L]
        seg += (float(rng.uniform(0.6, 1.5)) * tmpl[:len(seg)]).astype(np.complex64)
    start = 0
    for nb in blocks:
        nn = nb * ns
        two = bool(rng.integers(0, 2))
        a = ref.work(nn, x[start:start + nn + L], nitems_written=start, two_ports=two)
        b = ora.work(nn, x[start:start + nn + L], nitems_written=start, two_ports=two)
        for u, v in zip(a[:3], b[:3]):
            assert np.array_equal(np.asarray(u), np.asarray(v)), seed
        same_tags(a[3], b[3])
        start += nn


@pytest.mark.parametrize("seed", range(8))
def test_freqest_random_spectra(oracle, seed):
    rng = np.random.default_rng(5000 + seed)
    fftlen = int(rng.choice([64, 256, 1024, 4096]))
    nvec = int(rng.integers(1, 9))
    spec = (rng.standard_normal((nvec, fftlen)) + 1j * rng.standard_normal((nvec, fftlen))).astype(np.complex64)
    for v in range(nvec):
        r = rng.random()
        if r < 0.25:
            spec[v] = 0          # no energy: the previous estimate is repeated
        elif r < 0.5:
            spec[v, int(rng.integers(0, fftlen))] += 100 * fftlen
        elif r < 0.6:
            spec[v] = spec[v].real.astype(np.complex64)   # ties between mirrored bins
    a = R.freqest_work(spec, 48000.0, 9600, fftlen)
    b = oracle.freqest_work(spec, 48000.0, 9600, fftlen)[0]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("seed", range(10))
def test_chain_and_stream_random_configurations(oracle, seed):
    """the whole chain and its stream form on the reference's classes against the restated blocks:
    random template kind, threshold, mark delay, loop gain, fftlen, stages, record and piece sizes"""
    from gr_ais_b200 import synth
    from gr_ais_b200.ais_demod import preamble_template
    rng = np.random.default_rng(3000 + seed)
    tmpl = preamble_template(str(rng.choice(["north_star", "north_star", "intended", "reference"])))
    stages = int(rng.choice([oracle.STAGE_FREQSYNC | oracle.STAGE_AGC, oracle.STAGE_AGC, 0]))
    cfg = oracle.chain_cfg(fftlen=int(rng.choice([256, 1024])), threshold=float(rng.choice([0.5, 0.8, 0.9])),
                           mark_delay=int(rng.integers(0, 4)), gain=float(rng.choice([0.04, 0.1, 0.175])),
                           limit=float(rng.choice([0.01, 0.05])), stages=stages)
    n = int(rng.integers(6000, 20000))
    x, _ = synth.make_record(400 + seed, n=n, nbursts=int(rng.integers(1, 4)),
                             snr_db=float(rng.uniform(6.0, 25.0)), random_impairments=True)
    a = oracle.demod_chain(x, tmpl, cfg, blocks=R.blocks())
    b = oracle.demod_chain(x, tmpl, cfg)
    assert np.array_equal(a["bits"], b["bits"])
    same_tags(a["tags"], b["tags"])
    if stages & oracle.STAGE_AGC:   # the stream form needs the reference's AGC window
        sa, sb = oracle.DemodStream(tmpl, cfg, blocks=R.blocks()), oracle.DemodStream(tmpl, cfg)
        cuts = np.unique(np.concatenate([[0, n], rng.integers(0, n, int(rng.integers(1, 7)))]))
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            (ab, at), (bb, bt) = sa.work(x[lo:hi]), sb.work(x[lo:hi])
            assert np.array_equal(ab, bb), (seed, lo, hi)
            same_tags(at, bt)
