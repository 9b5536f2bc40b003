"""ncu target: a few launches of the assembly / update kernels on one workload (plain launches, no graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
P = scene.make_scene(cfg); s = solver.LBASolver(0)
s.upload(P, abi.Options(abi.PROFILE_G, 0))
print("assemble ms", s.time_kernel(0, 2), "update ms", s.time_kernel(2, 2))
