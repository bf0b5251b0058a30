"""The CUDA fast path of gr::fxpt::float_to_fixed divides by pi with a reciprocal and two fmas;
this checks, exhaustively, that it is the IEEE quotient for every float in [0, pi_f]."""
import os
import subprocess


def test_reciprocal_division_by_pi_is_exact(tmp_path):
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exhaustive_div.c")
    exe = str(tmp_path / "exhaustive_div")
    subprocess.run(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", src, "-o", exe, "-lm"],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    total, bad = (int(v) for v in r.stdout.split())
    assert r.returncode == 0 and bad == 0
    assert total == 1078530012   # every float from +0 to pi_f inclusive
