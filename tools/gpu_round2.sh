#!/bin/bash
# One GPU session (round 2, second half): parity tests, smoke, bench + reference arm, then (arg "ncu") the launch list of the bench command
# and full captures of the dense Cholesky's trailing update.  Every step under its own timeout.  $2 = snapshot tag (default r02q).
T=${2:-r02q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_ref.json 2>/dev/null
for f in pytest.log smoke.log; do echo "== $f"; tail -c 1500 gpurun_out/${T}_$f; echo; done
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("C2 ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print("roofline_largest", {k: d["roofline_largest"][k] for k in ("ms_per_launch", "frac", "frac_fp64", "update_kernel_ms_per_launch", "solve_ms_per_trial")})
dc = d.get("dense_cholesky") or {}
print("dense", {k: dc.get(k) for k in ("ms_per_solve", "achieved", "frac")}, "loop closure", {k: (dc.get("loop_closure_window") or {}).get(k) for k in ("ms_per_lba", "lm_trials", "stage_ms_per_lba")})
print("configs", {k: v.get("ms_per_lba") for k, v in d.get("configs", {}).items()})
PY
if [ "$1" = "ncu" ]; then
  PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 5 > gpurun_out/plain.log 2>&1 &&
  PLBA_FORCE_DENSE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_syrk_dmma" -s 3 -c 1 -o gpurun_out/${T}_dense_syrk python tools/solve_only.py 5 > gpurun_out/ncu_syrk.log 2>&1
  tail -n 2 gpurun_out/ncu_syrk.log
  timeout 300 python tools/prof_assemble.py 5 > gpurun_out/pa5.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_w|k_update_w" -s 2 -c 2 -o gpurun_out/${T}_c5 python tools/prof_assemble.py 5 > gpurun_out/pa5_ncu.log 2>&1
  timeout 300 python tools/prof_run.py 2 > gpurun_out/pa2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_assemble|k_solve_small|k_update" -s 3 -c 3 -o gpurun_out/${T}_c2 python tools/prof_run.py 2 > gpurun_out/pa2_ncu.log 2>&1
  tail -n 1 gpurun_out/pa5_ncu.log; tail -n 1 gpurun_out/pa2_ncu.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-largest --no-configs > gpurun_out/plain2.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_C2_G.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-largest --no-configs > gpurun_out/ncu_l.log 2>&1
  tail -n 1 gpurun_out/ncu_l.log | cut -c1-300
fi
