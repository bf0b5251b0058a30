import sys, numpy as np
sys.path.insert(0, ".")
from gr_ais_b200 import blocks, synth, binding as B
from oracle import oracle as O
rng = np.random.default_rng(4)
x = synth.gmsk_modulate(rng.integers(0, 2, 800)).astype(np.complex64)
x += (0.05 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)
blk = blocks.msk_timing_recovery_cc.make(5.0, 0.04, 0.01, 1)
ref = O.MskBlock(5.0, 0.04, 0.01, 1)
nout = 1000
out = np.zeros((1, nout), np.complex64); err = np.zeros((1, nout), np.float32); mu = np.zeros((1, nout), np.float32)
k = blk.general_work(nout, [len(x)], [x], [out, err, mu])
r_out, r_err, r_mu, r_cons = ref.general_work(nout, x)
print("k", k, len(r_out), "sym equal", np.array_equal(out[0,:k], r_out), "err equal", np.array_equal(err[0,:k], r_err), "mu equal", np.array_equal(mu[0,:k], r_mu))
bad = np.nonzero(err[0,:k] != r_err)[0]
print("bad", len(bad), bad[:10])
for i in bad[:5]:
    print(i, err[0,i], r_err[i], err[0,i].view(np.uint32) if hasattr(err[0,i],'view') else None, np.float32(r_err[i]).view(np.uint32), "sym", out[0,i], r_out[i])
