mkdir -p gpurun_out
PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 5 > gpurun_out/r02j_dense_plain.log 2>&1
cat gpurun_out/r02j_dense_plain.log | tail -3
PLBA_FORCE_DENSE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02j_launches_dense_C5.csv python tools/solve_only.py 5 > gpurun_out/r02j_dense_ncu.log 2>&1
tail -2 gpurun_out/r02j_dense_ncu.log
