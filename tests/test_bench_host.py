"""Host logic of bench.py that needs no GPU: the record plan every rank / the CPU arm builds its
channels from, the BASELINE configuration naming, and the reference arm's one JSON line (rank 0
only under a multi-rank launch)."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    a = dict(seconds=0.25, snr_db=20.0, records="independent", channels=0, workload="chain", template="north_star")
    a.update(kw)
    return argparse.Namespace(**a)


def test_record_plan_is_a_function_of_the_global_channel_number():
    n = 12000
    whole = bench.record_plan(_args(), n, 96)
    parts = [bench.record_plan(_args(), n, 48, first_channel=48 * r) for r in range(2)]
    assert np.array_equal(np.concatenate([p[1] for p in parts]), whole[1])
    assert np.array_equal(np.concatenate([p[2] for p in parts]), whole[2])
    assert np.array_equal(parts[0][0], whole[0]) and whole[0].shape == (bench.POOL, n)
    assert len(set(whole[2].tolist())) > 80           # independent rotations
    rows = bench.host_rows(whole, 3, 6)
    for k, c in enumerate(range(3, 6)):
        assert np.array_equal(rows[k], np.roll(whole[0][whole[1][c]], int(whole[2][c])))
    co = bench.record_plan(_args(records="coherent"), n, 8, first_channel=4)
    assert co[0].shape == (1, n) and list(co[2]) == [(16 * c) % n for c in range(4, 12)]


def test_defaults_are_the_baseline_configurations():
    assert bench.default_channels(_args(), 1) == 65536
    assert bench.default_channels(_args(), 8) == 32768
    assert bench.default_channels(_args(channels=4096), 1) == 4096
    assert bench.workload_name(_args(), 48000, 65536, 1).startswith("BASELINE configs[2]: 65536 channels/GPU")
    assert bench.workload_name(_args(), 48000, 32768, 8).startswith("BASELINE configs[3] (8 of 8 GPUs")
    assert bench.workload_name(_args(workload="corr_msk"), 48000, 4096, 1).startswith("BASELINE configs[1]")
    assert "BASELINE" not in bench.workload_name(_args(), 48000, 1000, 1)


def test_reference_arm_prints_one_line_on_rank_0_only():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--seconds", "0.2", "--cpu-channels", "4"]
    outs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", OMP_NUM_THREADS="2")
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append(r.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["unit"] == "channels/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference+shim", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "32768 channels/GPU" in line["config"]["workload"]
