# development aid: solver / update overlap of small single windows (PLBA_NO_OVERLAP=1 = plain chain): parity first, then C2 / C1 timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 or config2 or small_window or golden or edge_cases or graph" 2>&1 | tail -3
for i in 1 2; do
timeout 120 python tools/ab_bench.py pl_slam_plucker_b200/libplba.so 2 1 2>&1 | tail -2
PLBA_NO_OVERLAP=1 timeout 120 python tools/ab_bench.py pl_slam_plucker_b200/libplba.so 2 1 2>&1 | tail -2
done
