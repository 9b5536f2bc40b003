# development aid: dense tiled Cholesky at n = 12 000 (config 5 forced onto the dense path): parity tests, in-situ time of the default
# route against PLBA_DENSE_K1=1 (one trailing update per panel) and PLBA_NO_LOOKAHEAD=1, then the ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_window or loop_closure or randomised_large or config5" 2>&1 | tail -5
for i in 1 2; do
PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 5 2>&1 | tail -1
PLBA_DENSE_K1=1 PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 5 2>&1 | tail -1
done
PLBA_NO_LOOKAHEAD=1 PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 5 2>&1 | tail -1
PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 4 2>&1 | tail -1
PLBA_NO_LOOKAHEAD=1 PLBA_FORCE_DENSE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${1:-r02k}_launches_dense_C5.csv python tools/solve_only.py 5 > gpurun_out/dense_ncu.log 2>&1
tail -2 gpurun_out/dense_ncu.log
