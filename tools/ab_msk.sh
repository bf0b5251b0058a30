#!/bin/bash
# A/B of k_msk variants (env switches in launch_msk) at several batch sizes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for ch in 4096 16384 32768 65536; do
  for v in "" ; do
    echo "== channels $ch variant [$v]"
    env $v python bench.py --channels $ch --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-coherent --no-sc16 2>/dev/null | head -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})"
  done
done
