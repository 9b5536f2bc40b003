#!/bin/bash
# A/B of two builds of the library on ONE box: pl_slam_plucker_b200/libplba.so (new) against pl_slam_plucker_b200/libplba_base.so (built from an
# older tree with `python -m pl_slam_plucker_b200.build --out=…/libplba_base.so`), alternating, C2 bench line only.  $1 = tag, $2 = pairs (default 2).
T=${1:-ab}; N=${2:-2}; D=pl_slam_plucker_b200
cp $D/libplba.so /tmp/libplba_new.so
for i in $(seq 1 $N); do
  cp /tmp/libplba_new.so $D/libplba.so
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-largest --no-configs > gpurun_out/${T}_new_$i.json 2>/dev/null
  cp $D/libplba_base.so $D/libplba.so
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-largest --no-configs > gpurun_out/${T}_base_$i.json 2>/dev/null
done
cp /tmp/libplba_new.so $D/libplba.so
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "resident ms", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d["e2e"]["breakdown_ms"])
PY
