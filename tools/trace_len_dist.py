"""Run-to-run spread of the LM trace of one small window (atomics make the summation order run-dependent; a stalled LM decides on rounding):
python tools/trace_len_dist.py <lib.so> [repeats]   — prints the trace lengths / trial counts seen and the worst pose deviation from the first solve."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pl_slam_plucker_b200 import abi, scene, solver, _lib
lib = _lib.load(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
s = solver.LBASolver(0, lib=lib)
shapes = [dict(n_kf_free=21, n_kf_fixed=3, n_pt=1500, n_ls=400), dict(n_kf_free=20, n_kf_fixed=2, n_pt=8000, n_ls=2000), dict(n_kf_free=10, n_kf_fixed=2, n_pt=2000, n_ls=500)]
for si, (kw, seed) in enumerate(zip(shapes, (904, 902, 901))):
    P = scene.make_scene(1, seed=seed, **kw)
    for q in (0, 1):
        first = None; c = collections.Counter(); worst = 0.0
        for it in range(n):
            r = s.solve(P, abi.Options(abi.PROFILE_G, q))
            c[len(r.trace)] += 1
            if first is None: first = r
            else: worst = max(worst, float(np.abs(r.kf_T_wc - first.kf_T_wc).max()))
        print("shape", si, "quirks", q, "trace lengths", dict(sorted(c.items())), "worst pose deviation %.2e" % worst, flush=True)
