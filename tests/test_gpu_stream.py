"""GPU parity of the streamed ais_demod chain (state carried from call to call) against the
oracle's stream restatement, call by call and bit-exact."""
import numpy as np
import pytest

from gr_ais_b200 import binding as B
from gr_ais_b200 import synth
from gr_ais_b200.ais_demod import ais_demod

pytestmark = pytest.mark.gpu


def _records(channels, n, **kw):
    recs = [synth.make_record(c, n=n, **kw) for c in range(channels)]
    return np.stack([r[0] for r in recs]), [r[1] for r in recs]


def _cfg(oracle, d, stages, **kw):
    return oracle.chain_cfg(stages=stages, sample_rate=d.cfg.sample_rate, data_rate=d.cfg.data_rate,
                            fftlen=d.cfg.fftlen, sps=d.cfg.sps, gain=d.cfg.gain, limit=d.cfg.limit,
                            threshold=d.cfg.threshold, corr_chunk=d.cfg.corr_chunk, **kw)


def _run_stream(oracle, x, tmpl, pieces, stages=B.STAGE_FREQSYNC | B.STAGE_AGC, options=None, corr_chunk=0,
                threshold=0.9):
    """Feed x in the given piece sizes to both implementations; compare every call's output."""
    C, n = x.shape
    assert sum(pieces) == n
    d = ais_demod(options, channels=C, max_samples=max(max(pieces), 1), template=tmpl, stages=stages,
                  max_tags=2048, corr_chunk=corr_chunk, threshold=threshold)
    refs = [oracle.DemodStream(tmpl, _cfg(oracle, d, stages)) for _ in range(C)]
    all_bits = [[] for _ in range(C)]
    pos = 0
    for call, m in enumerate(pieces):
        bits, nbits, tags, ntags = d.stream_work(x[:, pos:pos + m])
        for c in range(C):
            rb, rt = refs[c].work(x[c, pos:pos + m])
            assert nbits[c] == len(rb), "call %d channel %d: %d bits, oracle %d" % (call, c, nbits[c], len(rb))
            assert np.array_equal(bits[c, :nbits[c]], rb), "call %d channel %d: bits differ" % (call, c)
            assert ntags[c] == len(rt), "call %d channel %d: %d tags, oracle %d" % (call, c, ntags[c], len(rt))
            for f in ("offset", "key", "port", "value"):
                assert np.array_equal(tags[c, :ntags[c]][f], rt[f]), (call, c, f)
            all_bits[c].append(bits[c, :nbits[c]].copy())
        pos += m
    pend = d.stream_pending()
    d.close()
    return [np.concatenate(b) if b else np.zeros(0, np.uint8) for b in all_bits], pend


def test_stream_equals_oracle_stream_odd_pieces(oracle, templates):
    x, truth = _records(5, 40000, nbursts=4, snr_db=25)
    pieces = [1, 1023, 4097, 0, 137, 6000, 12345, 3, 999, 15395]
    bits, pend = _run_stream(oracle, x, templates[120], pieces)
    found = sum(sum(synth.payloads_found(bits[c], truth[c])) for c in range(5))
    assert found >= 15, found
    assert pend[0] == 40000 % 1024


def test_stream_equals_batch_when_fed_whole(oracle, templates):
    """one stream call from fresh state produces the batch call's bits"""
    x, _ = _records(4, 16384, nbursts=3, snr_db=25)
    d = ais_demod(channels=4, max_samples=16384, template=templates[120], max_tags=1024)
    b0, n0, t0, nt0 = d.work(x)
    b1, n1, t1, nt1 = d.stream_work(x)
    assert np.array_equal(n0, n1) and np.array_equal(nt0, nt1)
    for c in range(4):
        assert np.array_equal(b0[c, :n0[c]], b1[c, :n1[c]])
        assert np.array_equal(t0[c, :nt0[c]], t1[c, :nt1[c]])
    # and a batch call afterwards still starts from fresh blocks
    b2, n2, _, _ = d.work(x)
    assert np.array_equal(n0, n2)
    for c in range(4):
        assert np.array_equal(b0[c, :n0[c]], b2[c, :n2[c]])
    d.close()


def test_stream_split_invariance_of_the_bitstream(oracle, templates):
    """a cut on a whole FFT vector, a whole corr_est multiple and a whole work chunk (1024, 137
    and 137*64 all divide 1024*137): the concatenated bitstream does not depend on the cut,
    except where a time_est tag falls into the 3*sps/2 items general_work leaves unconsumed."""
    n = 2 * 1024 * 137
    x, _ = _records(3, n, nbursts=12, snr_db=25)
    whole, _ = _run_stream(oracle, x, templates[120], [n], corr_chunk=137 * 64)
    cut, _ = _run_stream(oracle, x, templates[120], [n // 2, n // 2], corr_chunk=137 * 64)
    for c in range(3):
        k = min(len(whole[c]), len(cut[c]))
        assert abs(len(whole[c]) - len(cut[c])) <= 2
        assert np.mean(whole[c][:k] == cut[c][:k]) > 0.98


@pytest.mark.parametrize("L", [140, 1120])
def test_stream_other_templates(oracle, templates, L):
    x, _ = _records(3, 30000, nbursts=3, snr_db=25)
    _run_stream(oracle, x, templates[L], [5000, 7001, 2999, 15000])


def test_stream_without_freq_sync(oracle, templates):
    x, _ = _records(3, 20000, nbursts=3, snr_db=25)
    _run_stream(oracle, x, templates[120], [333, 10000, 1, 9666], stages=B.STAGE_AGC)


def test_stream_small_fft_and_chunk(oracle, templates):
    x, _ = _records(3, 20000, nbursts=3, snr_db=25)
    _run_stream(oracle, x, templates[120], [2500] * 8, options={"fftlen": 256}, corr_chunk=137 * 5)


def test_stream_reset_restarts(oracle, templates):
    x, _ = _records(2, 12000, nbursts=2, snr_db=25)
    d = ais_demod(channels=2, max_samples=12000, template=templates[120], max_tags=1024)
    a = d.stream_work(x)
    d.stream_work(x[:, :5000])
    d.stream_reset()
    b = d.stream_work(x)
    assert np.array_equal(a[1], b[1])
    for c in range(2):
        assert np.array_equal(a[0][c, :a[1][c]], b[0][c, :b[1][c]])
    assert np.array_equal(a[3], b[3])
    d.close()


def test_stream_needs_the_agc_stage(templates):
    d = ais_demod(channels=1, max_samples=4096, template=templates[120], stages=B.STAGE_FREQSYNC)
    with pytest.raises(B.B200AisError):
        d.stream_work(np.zeros((1, 4096), np.complex64))
    d.close()


def test_stream_many_channels_dev(oracle, templates):
    """device-resident variant, enough channels for several warps of the per-channel kernels"""
    torch = pytest.importorskip("torch")
    C, n = 96, 24000
    x, _ = _records(C, n, nbursts=3, snr_db=22)
    d = ais_demod(channels=C, max_samples=8192, template=templates[120], max_tags=1024)
    refs = [oracle.DemodStream(templates[120], _cfg(oracle, d, B.STAGE_FREQSYNC | B.STAGE_AGC)) for _ in range(C)]
    pieces = [8192, 7000, 8192, 616]
    mb = d.stream_max_bits(8192)
    bits = torch.zeros((C, mb), dtype=torch.uint8, device="cuda")
    nbits = torch.zeros(C, dtype=torch.int32, device="cuda")
    pos = 0
    for m in pieces:
        xd = torch.from_numpy(np.ascontiguousarray(x[:, pos:pos + m])).cuda()
        d.stream_work_dev(xd.data_ptr(), m, bits.data_ptr(), mb, nbits.data_ptr(),
                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        d.status()
        hb, hn = bits.cpu().numpy(), nbits.cpu().numpy()
        for c in range(C):
            rb, _ = refs[c].work(x[c, pos:pos + m])
            assert hn[c] == len(rb), (pos, c)
            assert np.array_equal(hb[c, :hn[c]], rb), (pos, c)
        pos += m
    d.close()


def test_stream_set_symbols_hot_swap(oracle, templates):
    """set_symbols between stream calls: taps swapped verbatim, threshold and filter tail kept
    (lib/corr_est_cc_impl.cc:132-162); same call sequence on the oracle stream"""
    x, _ = _records(4, 30000, nbursts=4, snr_db=25)
    t0 = templates[120]
    # the taps the constructor would have built from a time-shifted copy of the preamble, handed
    # over the way set_symbols stores them (no conjugate / reversal)
    t1 = np.ascontiguousarray(np.conj(np.roll(t0, 3))[::-1])
    d = ais_demod(channels=4, max_samples=12000, template=t0, max_tags=2048)
    refs = [oracle.DemodStream(t0, _cfg(oracle, d, B.STAGE_FREQSYNC | B.STAGE_AGC)) for _ in range(4)]
    pos = 0
    for call, m in enumerate([9000, 12000, 9000]):
        if call == 1:
            d.set_symbols(t1)
            for r in refs:
                r.set_symbols(t1)
        bits, nbits, tags, ntags = d.stream_work(x[:, pos:pos + m])
        for c in range(4):
            rb, rt = refs[c].work(x[c, pos:pos + m])
            assert nbits[c] == len(rb) and np.array_equal(bits[c, :nbits[c]], rb), (call, c)
            assert ntags[c] == len(rt), (call, c)
            for f in ("offset", "key", "port", "value"):
                assert np.array_equal(tags[c, :ntags[c]][f], rt[f]), (call, c, f)
        pos += m
    with pytest.raises(B.B200AisError):
        d.set_symbols(templates[140])
    d.close()
