import os, sys, time
sys.path.insert(0, os.getcwd())
from pl_slam_plucker_b200 import abi, scene, solver
s = solver.LBASolver(0)
P = scene.make_scene(5); opt = abi.Options(abi.PROFILE_G, 0, iters_stage1=1, iters_stage2=0)
for i in range(3):
    t = time.time(); s.upload(P, opt); print("upload wall ms", round((time.time() - t) * 1e3, 1), flush=True)
r = None
