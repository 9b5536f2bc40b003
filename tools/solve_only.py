"""Time of one reduced-system solve on a resident upload: python tools/solve_only.py <config> [lib.so]   (PLBA_FORCE_DENSE=1: dense path)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver, _lib
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lib = _lib.load(sys.argv[2]) if len(sys.argv) > 2 else None
P = scene.make_scene(cfg); s = solver.LBASolver(0, lib=lib)
s.upload(P, abi.Options(abi.PROFILE_G, 1))
print(s.time_kernel(1, 1), s.time_kernel(1, 3))
