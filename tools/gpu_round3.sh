#!/bin/bash
# ncu evidence of the third session of round 2 (the tests / bench of the same tree ran in the preceding call): launch list of the bench command,
# full captures of the upload's gather (k_reset) and of k_export at config 5, and of the three hot kernels at config 2.  $1 = tag.
T=${1:-r02zh}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-largest --no-configs > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_C2_G.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-largest --no-configs > gpurun_out/ncu_l.log 2>&1
tail -n 1 gpurun_out/ncu_l.log | cut -c1-300
timeout 300 python tools/prof_assemble.py 5 > gpurun_out/pa5.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_reset|k_assemble_w|k_update_w" -c 4 -o gpurun_out/${T}_c5 python tools/prof_assemble.py 5 > gpurun_out/pa5_ncu.log 2>&1
tail -n 1 gpurun_out/pa5_ncu.log
timeout 300 python tools/prof_run.py 2 > gpurun_out/pa2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_reset|k_assemble|k_solve_small|k_update|k_export" -c 6 -o gpurun_out/${T}_c2 python tools/prof_run.py 2 > gpurun_out/pa2_ncu.log 2>&1
tail -n 1 gpurun_out/pa2_ncu.log
ls -la gpurun_out/${T}_*
