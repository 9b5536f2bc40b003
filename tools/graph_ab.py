"""Whole-LBA wall time of the CUDA-graph path on C1 / C2 (A/B of graph layouts): python tools/graph_ab.py [lib.so]"""
import os, sys, time, json
sys.path.insert(0, '/root/repo')
from pl_slam_plucker_b200 import abi, scene, solver, _lib
import torch
s = solver.LBASolver(0, lib=_lib.load(sys.argv[1]) if len(sys.argv) > 1 else None)
for cfg in (1, 2):
    P = scene.make_scene(cfg); s.upload(P, abi.Options(abi.PROFILE_G, 0))
    for _ in range(5): s.reset(); s.run()
    ts = []
    for _ in range(30):
        s.reset(); torch.cuda.synchronize(); t = time.perf_counter(); s.run(); ts.append(time.perf_counter() - t)
    print(json.dumps({"cfg": cfg, "run_ms_min": round(1e3 * min(ts), 4), "run_ms_med": round(1e3 * sorted(ts)[15], 4), "trials": s.timing()["n_trials_run"], "launches": s.timing()["n_launches_run"]}))
