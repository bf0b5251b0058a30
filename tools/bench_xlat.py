#!/usr/bin/env python3
"""Device-resident timing of the channeliser (k_rot_phase + k_xlat_fir) on one B200.

    python tools/bench_xlat.py [--sources 2048] [--rate 250e3] [--seconds 1.0] [--freqs -25e3,25e3]
Prints one JSON line: channel-seconds of output per second, input GB/s, fma TFLOP/s."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gr_ais_b200 import binding as B  # noqa: E402
from gr_ais_b200 import blocks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=2048)
    ap.add_argument("--rate", type=float, default=250e3)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--freqs", default="-25e3,25e3")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    freqs = [float(f) for f in a.freqs.split(",")]
    B.set_device(0)
    torch.cuda.set_device(0)
    taps = blocks.firdes_low_pass(1.0, a.rate, 11e3, 1e3)
    D = int(a.rate / 48000)
    nout = int(a.seconds * a.rate / D)
    nin = len(taps) - 1 + nout * D
    stride = (nin + 1) & ~1
    blk = blocks.freq_xlating_fir_filter_ccf(D, taps, freqs, a.rate, sources=a.sources)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((a.sources, stride, 2), device="cuda", dtype=torch.float32, generator=g)
    ostride = (nout + 1) & ~1
    y = torch.empty((a.sources * len(freqs), ostride, 2), device="cuda", dtype=torch.float32)
    s = torch.cuda.current_stream().cuda_stream
    L = B.lib()

    def step():
        B.check(L.b200ais_xlat_work_dev(blk._h, nout, x.data_ptr(), stride, y.data_ptr(), ostride, s))

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    chans = a.sources * len(freqs)
    out_rate = a.rate / D
    passes = len(freqs) - (1 if len(freqs) == 2 and freqs[0] == -freqs[1] else 0)
    fma = 4.0 * len(taps) * nout * a.sources * passes
    print(json.dumps({
        "kernel": "k_xlat_fir", "sources": a.sources, "rate": a.rate, "decimation": D,
        "ntaps": len(taps), "freqs": freqs, "noutput": nout, "ms_per_step": ms,
        "channels_per_s": chans * (nout / out_rate) / (ms * 1e-3),
        "input_GBps": a.sources * nin * 8 / (ms * 1e-3) / 1e9,
        "fp32_fma_TFLOPs": 2 * fma / (ms * 1e-3) / 1e12,
        "input_bytes": a.sources * nin * 8}))


if __name__ == "__main__":
    main()
