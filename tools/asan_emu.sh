#!/bin/bash
# Kernel indexing under AddressSanitizer: the kernel sources compiled for the host emulation (-DPLBA_HOST_EMU; shared memory is a
# heap block of the exact size per launch) run small cases of every kernel family.  Test tooling only.
set -e
cd "$(dirname "$0")/.."
/usr/bin/g++ -O1 -g -std=c++17 -fPIC -fopenmp -shared -DPLBA_HOST_EMU -fsanitize=address -fno-omit-frame-pointer -x c++ \
    pl_slam_plucker_b200/csrc/plba_api.cu pl_slam_plucker_b200/csrc/scene_gen.cpp -o /tmp/libplba_emu_asan.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 PLBA_ASAN_SO=/tmp/libplba_emu_asan.so python tools/asan_run.py
