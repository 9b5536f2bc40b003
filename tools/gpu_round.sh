#!/bin/bash
# One GPU session: parity tests, smoke, bench (graph and host-loop), reference arm.  Every step under its own timeout.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1
PLBA_NO_GRAPH=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
for f in pytest_gpu smoke bench bench_nograph bench_ref; do echo "== $f"; tail -c 1500 gpurun_out/$f.log; echo; done
