"""world_size-2 gloo test of the only multi-GPU exchange in the path: rank 0 broadcasts the
preamble template, every rank takes its own contiguous slice of channels, timings reduce
with MAX.  Runs on CPU (the GPU box uses the same code over NCCL)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from gr_ais_b200 import sharding
    from gr_ais_b200.ais_demod import preamble_template
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    tmpl = sharding.broadcast_template(preamble_template("north_star") if rank == 0 else None, src=0)
    lo, hi = sharding.partition(4097, world, rank)
    t = sharding.max_over_ranks(1.0 + rank)
    print(json.dumps({"rank": rank, "n": int(len(tmpl)), "sum": float(np.abs(tmpl).sum()),
                      "first": [float(tmpl[0].real), float(tmpl[0].imag)], "lo": lo, "hi": hi, "t": t}))
    dist.destroy_process_group()
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_template_broadcast_and_partition_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2",
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        outs.append(out.strip().splitlines()[-1])
    import json
    r = sorted((json.loads(o) for o in outs), key=lambda d: d["rank"])
    assert r[0]["n"] == r[1]["n"] == 120
    assert r[0]["sum"] == r[1]["sum"] and r[0]["first"] == r[1]["first"]
    assert (r[0]["lo"], r[0]["hi"], r[1]["lo"], r[1]["hi"]) == (0, 2048, 2048, 4097)
    assert r[0]["t"] == r[1]["t"] == 2.0
