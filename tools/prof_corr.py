#!/usr/bin/env python3
"""Runs a few chain steps on a small batch so that `ncu -k regex:<kernel>` can capture one kernel.
    ncu --set full --clock-control none --import-source on -k regex:k_corr_fft -s 2 -c 1 -o gpurun_out/prof python tools/prof_corr.py [channels]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gr_ais_b200 import binding as B  # noqa: E402
from gr_ais_b200 import synth  # noqa: E402
from gr_ais_b200.ais_demod import ais_demod, preamble_template  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = 48000
B.set_device(0)
base, _ = synth.make_record(0, n=n, nbursts=4, snr_db=20.0)
x = torch.from_numpy(np.stack([np.roll(base, 16 * c) for c in range(256)]).view(np.float32).reshape(256, n, 2)).cuda()
x = x.repeat(C // 256, 1, 1).contiguous()
d = ais_demod(channels=C, max_samples=n, template=preamble_template(sys.argv[2] if len(sys.argv) > 2 else "north_star"))
mb = d.max_bits(n)
bits = torch.zeros((C, mb), dtype=torch.uint8, device="cuda")
nbits = torch.zeros(C, dtype=torch.int32, device="cuda")
for _ in range(3):
    d.work_dev(x.data_ptr(), n, bits.data_ptr(), mb, nbits.data_ptr(), None, None, None)
torch.cuda.synchronize()
d.status()
print("ok", int(nbits[0]))
