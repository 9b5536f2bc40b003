#!/bin/bash
# One GPU session: parity tests, smoke, bench (default routing; the C5 roofline runs on the warp kernels), reference arm, then the
# ncu launch list of the same bench command and full captures of the dominant kernels.  Every step under its own timeout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 python tools/batch_bench.py 1024 > gpurun_out/batch.log 2>&1
for f in pytest_gpu smoke bench bench_ref batch; do echo "== $f"; tail -c 3000 gpurun_out/$f.log; echo; done
if [ "$1" = "ncu" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-largest > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01k_C2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-largest > gpurun_out/ncu_l.log 2>&1
  timeout 300 python tools/prof_assemble.py 5 > gpurun_out/pa5.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_w|k_update_w" -s 2 -c 2 -o gpurun_out/prof_r01k_c5 python tools/prof_assemble.py 5 > gpurun_out/pa5_ncu.log 2>&1
  timeout 300 python tools/prof_run.py 2 > gpurun_out/pa2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_assemble|k_solve_small|k_update" -s 3 -c 3 -o gpurun_out/prof_r01k_c2 python tools/prof_run.py 2 > gpurun_out/pa2_ncu.log 2>&1
  tail -n 2 gpurun_out/pa5_ncu.log; tail -n 2 gpurun_out/pa2_ncu.log
fi
