"""b200ais_demod_work_sc16: the chain fed with interleaved int16 I/Q (the SDR wire format; the
reference's UHD / osmosdr sources convert it to gr_complex on the host, python/radio.py:151-203).
The device conversion is float(v) * scale per component, so the result must equal the oracle run
on the same quantised samples converted on the host, bit for bit."""
import numpy as np
import pytest

from gr_ais_b200 import binding as B
from gr_ais_b200 import synth
from gr_ais_b200.ais_demod import ais_demod, preamble_template

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stages", ["chain", "corr_msk"])
@pytest.mark.parametrize("scale", [1.0 / 8192.0, 3.0517578125e-05, 1.3e-4])
def test_sc16_equals_oracle_on_the_quantised_samples(oracle, stages, scale):
    C, n = 6, 16384 + 37
    tmpl = preamble_template("north_star")
    x = np.stack([synth.make_record(90 + c, n=n, nbursts=2, snr_db=18.0, random_impairments=True)[0]
                  for c in range(C)])
    q = np.empty((C, n, 2), np.int16)
    q[..., 0] = np.clip(np.rint(x.real / scale), -32768, 32767).astype(np.int16)
    q[..., 1] = np.clip(np.rint(x.imag / scale), -32768, 32767).astype(np.int16)
    xf = (q[..., 0].astype(np.float32) * np.float32(scale)) + 1j * (q[..., 1].astype(np.float32) * np.float32(scale))
    xf = xf.astype(np.complex64)
    st = (B.STAGE_FREQSYNC | B.STAGE_AGC) if stages == "chain" else 0
    d = ais_demod(channels=C, max_samples=n, template=tmpl, stages=st)
    bits, nbits, tags, ntags = d.work_sc16(q, scale)
    b2, nb2, t2, nt2 = d.work(xf)            # the same samples through the float entry point
    d.close()
    cfg = oracle.chain_cfg(stages=(oracle.STAGE_FREQSYNC | oracle.STAGE_AGC) if stages == "chain" else 0)
    for c in range(C):
        r = oracle.demod_chain(xf[c], tmpl, cfg)
        assert nbits[c] == len(r["bits"]) == nb2[c]
        assert np.array_equal(bits[c, :nbits[c]], r["bits"])
        assert np.array_equal(b2[c, :nb2[c]], r["bits"])
        for f in ("offset", "key", "port", "value"):
            assert np.array_equal(tags[c, :ntags[c]][f], r["tags"][f]), f
    if stages == "chain":      # (without the AGC a unit-amplitude burst stays below 0.9 L^2)
        assert int(ntags.sum()) > 0


def test_sc16_argument_errors():
    d = ais_demod(channels=2, max_samples=4096, template=preamble_template("north_star"))
    with pytest.raises(ValueError):
        d.work_sc16(np.zeros((3, 4096, 2), np.int16), 1.0)
    with pytest.raises(B.B200AisError):
        d.work_sc16(np.zeros((2, 5000, 2), np.int16), 1.0)
    d.close()
