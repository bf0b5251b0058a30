#!/bin/bash
# Scratch A/B driver for gpurun: bench lines for "VAR=VAL ..." variants (SKIP_TESTS=1 skips pytest).
#   bash tools/ab_run.sh "16384 65536" "" "B200AIS_MSK_NO_SLIDE=1"
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -4
chs="$1"; shift
for ch in $chs; do
  for v in "$@"; do
    echo "== channels $ch [$v]"
    env $v python bench.py --channels $ch --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-coherent --no-sc16 2>/dev/null | head -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d.get('serialized_ms_per_step'), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})"
  done
done
