import os, sys, time, json
sys.path.insert(0, os.getcwd())
from pl_slam_plucker_b200 import abi, scene, solver, _lib
lib = _lib.load(sys.argv[1])
s = solver.LBASolver(0, lib=lib)
for cfg in (4, 2):
    P = scene.make_scene(cfg); opt = abi.Options(abi.PROFILE_G, 0)
    for i in range(3):
        t = time.time(); r = s.solve(P, opt); dt = time.time() - t
        tm = s.timing()
        print(cfg, "solve wall ms", round(dt * 1e3, 2), {k: round(v, 3) for k, v in tm.items() if k.startswith("ms_")}, flush=True)
