#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (it runs the CPU oracle beside the GPU path, so it lives under tests/).
BASELINE.json configs[4]: SNR sweep, packet-detect rate of the GPU path vs the CPU oracle.

Per SNR point: `--channels` channels, independent noise per channel, random payloads, carrier
offset U(-500, 500) Hz, random phase, fractional delay U(0, 1) sample.
Reports, for the GPU path and (on a subset of channels) the oracle:
  detect rate = bursts with a corr_start tag within +-sps of where the preamble correlates
  crc rate    = bursts whose payload comes out of hdlc_deframer_bp(11, 64) with a good CRC
                (GPU path: b200ais_hdlc_work; oracle: its C restatement)
and checks the two paths agree bit for bit on the oracle subset.

    python tests/snr_sweep.py --channels 16384 --oracle-channels 1024 --procs 16 --snrs 0 2 4 6 8 10 12 14 16 18 20
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def expected_tag_offset(start, L, agc_delay=511, pulse_delay=20, ramp=8 * 5, template_delay=11.5):
    # training starts ramp samples into the burst; synth's full convolution delays it by
    # pulse_delay; the AGC by 511; corr_est's output 0 by L; the template's own Gaussian
    # filter start-up by ~11.5 samples; corr_start sits one item before the match
    return start + ramp + pulse_delay + agc_delay + L - template_delay - 1


def score(pdus, tags, ntags, truth, L, sps=5):
    det = crc = tot = 0
    for c in range(len(truth)):
        found = set(pdus[c])
        cs = tags[c, :ntags[c]]
        cs = cs[cs["key"] == 0]["offset"].astype(np.int64)
        for t in truth[c]:
            tot += 1
            crc += t["payload"] in found
            want = expected_tag_offset(t["start"], L)
            det += bool(len(cs)) and np.min(np.abs(cs - want)) <= 2 * sps
    return det, crc, tot


def _record(c, n, nbursts, snr):
    from gr_ais_b200 import synth
    return synth.make_record(c, n=n, nbursts=nbursts, snr_db=snr, random_impairments=True,
                             seed=synth.SEED + int(snr * 10))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--oracle-channels", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=0.5)
    ap.add_argument("--snrs", type=float, nargs="+", default=[0, 4, 8, 12, 16, 20])
    ap.add_argument("--threshold", type=float, default=0.9)
    ap.add_argument("--procs", type=int, default=0,
                    help="worker processes that synthesise the records (0 = in this process)")
    args = ap.parse_args()
    pool = None
    if args.procs > 1:  # forked before anything touches CUDA: the workers only run numpy
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(args.procs)
    from gr_ais_b200 import synth
    from gr_ais_b200.ais_demod import ais_demod, preamble_template
    from oracle import oracle as O

    # the CPU side runs the reference's own block classes (oracle/_ref) when they were built
    ref_blocks, cpu_kind = None, "oracle port"
    try:
        from oracle import ref as R
        if R.available():
            ref_blocks, cpu_kind = R.blocks(), "reference blocks (oracle/_ref) in the oracle's schedule"
    except Exception:
        pass
    n = int(args.seconds * 48000)
    tmpl = preamble_template("north_star")
    d = ais_demod(channels=args.channels, max_samples=n, template=tmpl, threshold=args.threshold,
                  max_tags=1024)
    from gr_ais_b200 import blocks
    deframer = blocks.hdlc_deframer_bp(11, 64, channels=args.channels)
    rows = []
    for snr in args.snrs:
        jobs = [(c, n, max(1, int(4 * args.seconds)), snr) for c in range(args.channels)]
        if pool is not None:
            recs = pool.starmap(_record, jobs, chunksize=64)
        else:
            recs = [_record(*j) for j in jobs]
        x = np.stack([r[0] for r in recs])
        truth = [r[1] for r in recs]
        bits, nbits, tags, ntags = d.work(x)
        deframer.reset()
        frames, nframes = deframer.work(bits, nbits, max_frames=32)
        det, crc, tot = score(deframer.pdus(frames, nframes), tags, ntags, truth, len(tmpl))
        k = min(args.oracle_channels, args.channels)
        ob, onb, ot, ont = O.demod_chain_batch(x[:k], tmpl, O.chain_cfg(threshold=args.threshold),
                                               max_tags=1024, blocks=ref_blocks)
        same = all(onb[c] == nbits[c] and np.array_equal(ob[c, :onb[c]], bits[c, :nbits[c]])
                   and ont[c] == ntags[c]
                   and np.array_equal(ot[c, :ont[c]]["offset"], tags[c, :ntags[c]]["offset"])
                   for c in range(k))
        opdus = [O.frames_payloads(O.HdlcDeframer(11, 64).work(ob[c, :onb[c]])) for c in range(k)]
        same = same and all(opdus[c] == deframer.pdus(frames, nframes)[c] for c in range(k))
        odet, ocrc, otot = score(opdus, ot.view(tags.dtype), ont, truth[:k], len(tmpl))
        rows.append(dict(snr_db=snr, bursts=tot, gpu_detect=det / tot, gpu_crc=crc / tot,
                         oracle_bursts=otot, oracle_detect=odet / otot, oracle_crc=ocrc / otot,
                         gpu_equals_oracle_on_subset=bool(same), cpu_path=cpu_kind,
                         channels=args.channels, oracle_channels=k, seconds=args.seconds))
        print(json.dumps(rows[-1]), flush=True)
    d.close()
    if pool is not None:
        pool.close()
    return rows


if __name__ == "__main__":
    main()
