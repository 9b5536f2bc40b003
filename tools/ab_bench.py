"""A/B micro-benchmark of the stage kernels (development aid): python tools/ab_bench.py <lib.so> [configs...]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver, _lib
lib = _lib.load(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else None
cfgs = [int(a) for a in sys.argv[2:]] or [2, 4]
s = solver.LBASolver(0, lib=lib)
for cfg in cfgs:
    t = time.time(); P = scene.make_scene(cfg); tg = time.time() - t
    opt = abi.Options(abi.PROFILE_G, 1)
    t = time.time(); s.upload(P, opt); tu = time.time() - t
    out = {"cfg": cfg, "n_obs": P.n_obs, "gen_s": round(tg, 2), "upload_s": round(tu, 3)}
    for name, which in (("assemble", 0), ("solve", 1), ("update", 2)):
        if which == 1 and P.n_free > 24: continue
        out[name + "_us"] = round(1e3 * s.time_kernel(which, 10 if cfg >= 4 else 50), 2)
    if cfg <= 4:
        ts = []
        for _ in range(3):
            s.reset(); t = time.time(); s.run(); ts.append(time.time() - t)
        tm = s.timing()
        out["run_ms"] = round(1e3 * min(ts), 3); out["trials"] = tm["n_trials_run"]; out["launches"] = tm["n_launches_run"]
    print(json.dumps(out), flush=True)
