import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P = scene.make_scene(cfg); s = solver.LBASolver(0)
s.upload(P, abi.Options(abi.PROFILE_G, 1))
print(s.time_kernel(1, 1))
