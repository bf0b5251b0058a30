"""Parity at the channel counts of BASELINE.json's configurations.

* configs[1]: 4096 channels through corr_est_cc + msk_timing_recovery_cc (+ the bit tail), and
  through the whole chain -- every 16th channel (256 of them) compared bit for bit with the
  oracle, and the per-channel symbol counts of ALL channels against a checksum property.
* the grid.z path: more than 65 535 channels in one launch (gr-ais_b200/csrc/internal.h
  channel_grid, which BASELINE configs[2]'s 65 536 channels need) on a short record, 272
  channels sampled across the seam at channel 65 535 and the grid.z split.
Channels are built from a small pool of seeded records, each with its own rotation, so that
identical inputs must give identical outputs (a size-independent property checked on every
channel) while a sampled subset is checked against the oracle."""
import numpy as np
import pytest

from gr_ais_b200 import binding as B
from gr_ais_b200 import synth
from gr_ais_b200.ais_demod import ais_demod, preamble_template

pytestmark = pytest.mark.gpu


def same_tags(got, want):
    assert len(got) == len(want)
    for f in ("offset", "key", "port", "value"):
        assert np.array_equal(got[f], want[f]), f


def build_batch(C, n, pool_size, seed0, snr_db=15.0):
    pool = [synth.make_record(seed0 + p, n=n, nbursts=max(1, n // 12000), snr_db=snr_db,
                              random_impairments=True)[0] for p in range(pool_size)]
    rng = np.random.default_rng(seed0)
    off = rng.integers(0, n, C)
    off[:pool_size] = 0                       # channel p < pool_size is record p unrotated
    pid = np.arange(C) % pool_size
    x = np.empty((C, n), np.complex64)
    for c in range(C):
        x[c] = np.roll(pool[pid[c]], int(off[c]))
    return x, pid, off


def check_against_oracle(oracle, x, tmpl, bits, nbits, tags, ntags, sample, stages):
    cfg = oracle.chain_cfg(stages=stages)
    for c in sample:
        r = oracle.demod_chain(x[c], tmpl, cfg)
        assert nbits[c] == len(r["bits"]), c
        assert np.array_equal(bits[c, :nbits[c]], r["bits"]), c
        same_tags(tags[c, :ntags[c]], r["tags"])


@pytest.mark.parametrize("workload", ["corr_msk", "chain"])
def test_configs1_4096_channels(oracle, workload):
    C, n = 4096, 12288
    tmpl = preamble_template("north_star")
    x, pid, off = build_batch(C, n, 32, 400)
    stages = 0 if workload == "corr_msk" else (B.STAGE_FREQSYNC | B.STAGE_AGC)
    d = ais_demod(channels=C, max_samples=n, template=tmpl, stages=stages)
    bits, nbits, tags, ntags = d.work(x)
    d.close()
    ostages = 0 if workload == "corr_msk" else (oracle.STAGE_FREQSYNC | oracle.STAGE_AGC)
    check_against_oracle(oracle, x, tmpl, bits, nbits, tags, ntags, range(0, C, 16), ostages)
    # every channel: equal inputs => equal outputs (channels c and c + 32k with the same rotation
    # do not exist by construction, so compare the unrotated copies planted at the far end)
    assert int(nbits.min()) > n // 5 - 64 and int(nbits.max()) < n // 5 + 64
    assert (ntags % 4 == 0).all()            # four tags per detection on port 0


def test_grid_z_split_above_65535_channels(oracle):
    C, n = 65600, 4096
    tmpl = preamble_template("north_star")
    x, pid, off = build_batch(C, n, 16, 500, snr_db=18.0)
    # plant unrotated copies of the pool on both sides of the grid.y limit and at the very end
    for base in (32768, 65520, 65536, C - 16):
        for p in range(16):
            x[base + p] = x[p]
    d = ais_demod(channels=C, max_samples=n, template=tmpl)
    bits, nbits, tags, ntags = d.work(x)
    d.close()
    sample = list(range(0, 16)) + list(range(32760, 32800)) + list(range(65500, 65600)) + \
        list(range(100, C, 569))
    assert len(sample) >= 256
    check_against_oracle(oracle, x, tmpl, bits, nbits, tags, ntags, sample,
                         oracle.STAGE_FREQSYNC | oracle.STAGE_AGC)
    for base in (32768, 65520, 65536, C - 16):
        for p in range(16):
            assert nbits[base + p] == nbits[p]
            assert np.array_equal(bits[base + p, :nbits[p]], bits[p, :nbits[p]])
            same_tags(tags[base + p, :ntags[base + p]], tags[p, :ntags[p]])


@pytest.mark.parametrize("C", [24000, 32000])
def test_timing_loop_geometries_between_the_configs(oracle, C):
    """The timing loop picks its shared-memory ring from the channel count: 24 000 channels run the
    96-sample ring with one warp per CTA (6 warps per SM), 32 000 (BASELINE configs[3]'s per-GPU
    batch) the same ring with 7 warps behind one table.  (4096 / 16 384 channels above take the
    128-sample ring, 65 600 the 48-sample one.)"""
    n = 4096
    tmpl = preamble_template("north_star")
    x, pid, off = build_batch(C, n, 16, 600, snr_db=18.0)
    d = ais_demod(channels=C, max_samples=n, template=tmpl)
    bits, nbits, tags, ntags = d.work(x)
    d.close()
    sample = list(range(0, 16)) + list(range(C - 40, C)) + list(range(7, C, C // 200))
    assert len(sample) >= 256
    check_against_oracle(oracle, x, tmpl, bits, nbits, tags, ntags, sample,
                         oracle.STAGE_FREQSYNC | oracle.STAGE_AGC)


@pytest.mark.parametrize("C,n", [(12320, 8192), (53280, 8192)])
def test_pipelined_submission_with_the_small_ring_equals_strict_calls(C, n):
    """enqueue_dev on >= 12 288 channels runs the timing loop with the 48-sample ring on the side
    stream (it shares the SMs with the next record's front kernels); work_dev on the same handle
    takes the 128-sample ring.  Above 53 248 channels enqueue_dev stays on the caller's stream.
    Record by record the bits, counts and tags must be the same."""
    torch = pytest.importorskip("torch")
    K = 3
    tmpl = preamble_template("north_star")
    recs = [build_batch(C, n, 16, 700 + k, snr_db=18.0)[0] for k in range(K)]
    d = ais_demod(channels=C, max_samples=n, template=tmpl)
    mb = d.max_bits(n)
    st = torch.cuda.Stream()
    xs = [torch.from_numpy(r.view(np.float32).reshape(C, n, 2)).cuda() for r in recs]

    def buffers():
        return (torch.zeros((C, mb), dtype=torch.uint8, device="cuda"),
                torch.zeros(C, dtype=torch.int32, device="cuda"),
                torch.zeros((C, d.max_tags, 24), dtype=torch.uint8, device="cuda"),
                torch.zeros(C, dtype=torch.int32, device="cuda"))
    ref = []
    for k in range(K):
        b = buffers()
        d.work_dev(xs[k].data_ptr(), n, b[0].data_ptr(), mb, b[1].data_ptr(), b[2].data_ptr(), b[3].data_ptr(),
                   st.cuda_stream)
        st.synchronize()
        d.status()
        ref.append([t.cpu().numpy() for t in b])
    outs = []
    for k in range(K):
        b = buffers()
        d.enqueue_dev(xs[k].data_ptr(), n, b[0].data_ptr(), mb, b[1].data_ptr(), b[2].data_ptr(), b[3].data_ptr(),
                      st.cuda_stream)
        outs.append(b)
    d.join(st.cuda_stream)
    st.synchronize()
    d.status()
    for k in range(K):
        bits, nb, tg, nt = (t.cpu().numpy() for t in outs[k])
        rb, rn, rt, rnt = ref[k]
        assert np.array_equal(nb, rn) and np.array_equal(nt, rnt), k
        assert int(nt.sum()) > 0
        for c in range(0, C, 7):
            assert np.array_equal(bits[c, :nb[c]], rb[c, :nb[c]]), (k, c)
            assert np.array_equal(tg[c, :nt[c]], rt[c, :nt[c]]), (k, c)
    d.close()


@pytest.mark.parametrize("env", [
    {"B200AIS_MSK_KIND": "2", "B200AIS_FUSE_TAIL_MIN_CH": "1"},
    {"B200AIS_MSK_KIND": "1", "B200AIS_FUSE_TAIL_MIN_CH": "1"},
    {"B200AIS_MSK_KIND": "0", "B200AIS_FUSE_TAIL_MIN_CH": "1"},
    {"B200AIS_MSK_KIND": "2", "B200AIS_NO_FUSE_TAIL": "1", "B200AIS_MSK_NO_PAIR_FETCH": "1"},
    {"B200AIS_MSK_NO_SLIDE": "1"},
])
def test_full_occupancy_code_paths_on_a_small_batch(env):
    """the timing-loop variants a large batch selects (48-sample ring, paired fetch, bits written by
    the loop) forced on the smoke batch in a fresh process (the switches are read once): bits,
    symbol counts and tags against the oracle"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, env=e,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-2000:]
