mkdir -p gpurun_out
for g in 2 3 4 6; do
echo "G=$g"; PLBA_DENSE_GROUP=$g timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_window or loop_closure or randomised_large_windows or config5" 2>&1 | tail -1
PLBA_DENSE_GROUP=$g PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 5 2>&1 | tail -1
PLBA_DENSE_GROUP=$g PLBA_FORCE_DENSE=1 timeout 300 python tools/solve_only.py 4 2>&1 | tail -1
done
