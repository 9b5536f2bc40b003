#!/bin/bash
# A/B of stage-kernel variants: bash tools/ab_round.sh "<lib|warp|chunk> ..." "<cfgs>"
mkdir -p gpurun_out
for lib in $1; do
  echo "== $lib"
  if [ "$lib" = "chunk" ]; then PLBA_FORCE_CHUNK=1 timeout 300 python tools/ab_bench.py - $2 2>&1 | tail -4
  elif [ "$lib" = "warp" ]; then PLBA_FORCE_CHUNK=2 timeout 300 python tools/ab_bench.py - $2 2>&1 | tail -4
  else PLBA_FORCE_CHUNK=2 timeout 300 python tools/ab_bench.py pl_slam_plucker_b200/$lib $2 2>&1 | tail -4; fi
done
