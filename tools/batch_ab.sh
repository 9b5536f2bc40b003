#!/bin/bash
for m in 1 2; do echo "== batch path $m"; PLBA_FORCE_CHUNK=$m timeout 600 python tools/batch_bench.py 1024 2>&1 | tail -1; done
