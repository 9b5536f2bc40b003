#!/bin/bash
# One GPU session: parity tests, smoke, bench (warp path and, for A/B, the CTA-chunk path), reference arm.  Every step under its own timeout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1
PLBA_FORCE_CHUNK=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
for f in pytest_gpu smoke bench bench_chunk bench_ref; do echo "== $f"; tail -c 2500 gpurun_out/$f.log; echo; done
