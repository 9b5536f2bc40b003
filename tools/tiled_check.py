"""GPU check of the tiled (large-window) solver: parity on n=180 / n=600 windows vs the oracle, timing on C4 and C5."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pl_slam_plucker_b200 import abi, scene, solver
from oracle import loader as orc
s = solver.LBASolver(0)
for nkf, npt, nls in ((30, 600, 150), (100, 6000, 1500), (200, 12000, 3000)):
    P = scene.make_scene(1, n_kf_free=nkf, n_kf_fixed=2, n_pt=npt, n_ls=nls, seed=7)
    opt = abi.Options(abi.PROFILE_G, 1)
    r = s.solve(P, opt); o = orc.solve(P, opt)
    n = min(len(r.trace), len(o.trace))
    print(json.dumps({"n": 6 * nkf, "chi_rel": float(np.max(np.abs(r.trace["chi"][:n] - o.trace["chi"][:n]) / o.trace["chi"][:n])),
                      "pose_diff": float(np.abs(r.kf_T_wc - o.kf_T_wc).max()), "trials": [int(r.n_trials), int(o.n_trials)]}), flush=True)
for cfg in [int(a) for a in sys.argv[1:]] or [4]:
    P = scene.make_scene(cfg)
    s.upload(P, abi.Options(abi.PROFILE_G, 1))
    s.time_kernel(1, 1)
    ms = s.time_kernel(1, 3)
    out = {"cfg": cfg, "n": 6 * P.n_free, "solve_ms": round(ms, 3), "gflops": round((6 * P.n_free) ** 3 / 3 / (ms * 1e-3) / 1e9, 1)}
    if cfg == 4:
        ts = []
        for _ in range(2):
            s.reset(); t = time.time(); s.run(); ts.append(time.time() - t)
        out["run_ms"] = round(1e3 * min(ts), 2); out["trials"] = s.timing()["n_trials_run"]
    print(json.dumps(out), flush=True)
