"""Quick on-GPU parity + timing probe (development aid): CUDA path through the C ABI vs the CPU oracle."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pl_slam_plucker_b200 import abi, scene, solver
from oracle import loader as orc

def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-12))) if a.size else 0.0

def compare(s, P, prof, q, label, oracle=True):
    opt = abi.Options(prof, q)
    t = time.time(); r = s.solve(P, opt); dt = time.time() - t
    out = {"label": label, "profile": prof, "quirks": q, "gpu_wall_s": round(dt, 4), "ntrace_gpu": int(len(r.trace)), "trials": int(r.n_trials), "timing": s.timing()}
    if oracle:
        t = time.time(); o = orc.solve(P, opt); out["oracle_s"] = round(time.time() - t, 3)
        n = min(len(o.trace), len(r.trace))
        out["ntrace_oracle"] = int(len(o.trace))
        fin = np.isfinite(o.trace['chi'][:n]) & np.isfinite(r.trace['chi'][:n])
        out["chi_rel_max"] = rel(o.trace['chi'][:n][fin], r.trace['chi'][:n][fin])
        out["accept_equal"] = bool((o.trace['accepted'][:n] == r.trace['accepted'][:n]).all())
        out["pose_diff"] = float(np.abs(o.kf_T_wc - r.kf_T_wc).max())
        out["pt_diff"] = float(np.abs(o.pt_xyz - r.pt_xyz).max()) if P.n_pt else 0.0
        if prof == abi.PROFILE_H_END:
            out["ls_diff"] = float(np.abs(o.ls_end - r.ls_end).max()) if P.n_ls else 0.0
        else:
            out["ls_diff"] = float(np.abs(o.ls_orth - r.ls_orth).max()) if P.n_ls else 0.0
        if prof == abi.PROFILE_G:
            out["flags_equal"] = bool((o.po_flags == r.po_flags).all() and (o.lo_flags == r.lo_flags).all())
    print(json.dumps(out), flush=True)
    return r

if __name__ == "__main__":
    s = solver.LBASolver(0)
    small = scene.make_scene(1, n_kf_free=4, n_kf_fixed=2, n_pt=60, n_ls=20)
    small_e = scene.make_scene(1, n_kf_free=4, n_kf_fixed=2, n_pt=60, n_ls=20, line_mode=1)
    for q in (0, 1):
        compare(s, small, abi.PROFILE_G, q, "small")
        compare(s, small_e, abi.PROFILE_H_END, q, "small")
        compare(s, small, abi.PROFILE_H_PLK, q, "small")
    c1 = scene.make_scene(1); c2 = scene.make_scene(2)
    for q in (0, 1):
        compare(s, c1, abi.PROFILE_G, q, "C1")
        compare(s, c2, abi.PROFILE_G, q, "C2")
    big = scene.make_scene(1, n_kf_free=30, n_kf_fixed=2, n_pt=600, n_ls=150, seed=7)
    compare(s, big, abi.PROFILE_G, 1, "tiled-solver n=180")
    # repeat timing, resident problem
    for name, P in (("C1", c1), ("C2", c2)):
        s.upload(P, abi.Options(abi.PROFILE_G, 1))
        ts = []
        for i in range(5):
            s.reset(); t = time.time(); s.run(); ts.append(time.time() - t)
        print(json.dumps({"resident_run": name, "wall_ms": [round(1e3 * x, 3) for x in ts], "timing": s.timing()}), flush=True)
    if "--c4" in sys.argv:
        c4 = scene.make_scene(4)
        compare(s, c4, abi.PROFILE_G, 1, "C4", oracle="--c4-oracle" in sys.argv)
