#!/usr/bin/env python3
"""gpurun_out/final/ (written on the GPU box by tools/collect_profiles.sh) -> profiles/r02_*:
bench lines, the ncu launch list, per-kernel ncu raw pages (CSV), a one-screen summary and the
SASS-level hot spots of every `ncu --set full` capture, compute-sanitizer logs."""
import glob
import io
import os
import shutil
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "final")
DST = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_sass_hot  # noqa: E402
import ncu_summary  # noqa: E402


def main():
    for f in sorted(glob.glob(os.path.join(SRC, "bench_*.json")) + glob.glob(os.path.join(SRC, "rx_*.json"))):
        if os.path.getsize(f):
            shutil.copy(f, os.path.join(DST, "r02_" + os.path.basename(f)))
    for f in ("launches_65536ch.csv", "sanitizer_memcheck_smoke.log", "sanitizer_memcheck_smoke_kind2_fused.log",
              "sanitizer_racecheck_smoke.log"):
        p = os.path.join(SRC, f)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, "r02_" + f))
    summary = io.StringIO()
    for rep in sorted(glob.glob(os.path.join(SRC, "*.ncu-rep"))):
        name = os.path.basename(rep)[:-len(".ncu-rep")]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        with open(os.path.join(DST, "r02_%s.csv" % name), "w") as fh:
            fh.write(raw)
        sys.argv = ["ncu_summary", rep]
        with redirect_stdout(summary):
            ncu_summary.main()
        hot = io.StringIO()
        sys.argv = ["ncu_sass_hot", rep, "24"]
        with redirect_stdout(hot):
            ncu_sass_hot.main()
        with open(os.path.join(DST, "r02_%s_sass_hot.txt" % name), "w") as fh:
            fh.write(hot.getvalue())
    with open(os.path.join(DST, "r02_ncu_summary.txt"), "w") as fh:
        fh.write(summary.getvalue().replace(SRC + "/", ""))
    print(summary.getvalue())


if __name__ == "__main__":
    main()
