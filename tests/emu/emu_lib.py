"""TEST TOOLING ONLY: builds and loads the host emulation of the CUDA kernels (csrc compiled with -DPLBA_HOST_EMU).

It exists so that kernel LOGIC (phase structure, indexing, controller) can be checked against the oracle on a box
without a GPU.  The package never loads it; `pl_slam_plucker_b200._lib.load()` only ever opens libplba.so.
"""
import os
import subprocess

from pl_slam_plucker_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
SO = os.path.join(HERE, "libplba_emu.so")
CSRC = os.path.join(ROOT, "pl_slam_plucker_b200", "csrc")
_L = None


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "plba.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-DPLBA_HOST_EMU", "-x", "c++",
                               os.path.join(CSRC, "plba_api.cu"), os.path.join(CSRC, "scene_gen.cpp"), "-o", SO])
    return SO


def load():
    global _L
    if _L is None:
        _L = _lib.declare(__import__("ctypes").CDLL(build()))
    return _L
