#!/usr/bin/env python3
"""Run bench.py over a few (channels, overlap) points and print one compact line each."""
import json
import subprocess
import sys

sys.argv_extra = [a for a in sys.argv[1:] if a.startswith("--")]
points = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:] if not a.startswith("--")] or [(4096, 4)]
for ch, ov in points:
    r = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--channels", str(ch),
                        "--overlap", str(ov)] + sys.argv_extra, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print("FAILED", ch, ov, r.stdout[-500:], r.stderr[-1500:])
        continue
    print("ch", ch, "ov", ov, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2),
          "serial_ms", round(d["serialized_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]),
          "e2e_ms", round(d["e2e"]["ms_per_step"], 1),
          {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, flush=True)
