"""Small driver for ncu captures: a few host-driven LM rounds of one workload (plain kernel launches, no graph)."""
import os, sys
os.environ.setdefault("PLBA_NO_GRAPH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = scene.make_scene(cfg)
s = solver.LBASolver(0)
s.upload(P, abi.Options(abi.PROFILE_G, 0))
for _ in range(steps):
    s.reset(); s.run()
print("ok", s.timing())
