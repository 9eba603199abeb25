#!/bin/bash
# One `ncu --set full` capture per hot kernel family of the bench step (run under gpurun; reports land in gpurun_out/).
# usage: tools/ncu_full.sh <tag> <kernel-regex> [launch-skip] [launch-count]
set -u
tag=$1; regex=$2; skip=${3:-0}; count=${4:-1}
mkdir -p gpurun_out
timeout 420 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c "$count" \
  -f -o "gpurun_out/ncu_$tag" python bench.py --profile --steps 1 --no-graph > "gpurun_out/ncu_$tag.log" 2>&1
echo "ncu $tag rc=$?"
