"""Measured deviation of the CUDA path from the CPU oracle per configuration (feeds the tolerances stated in tests/ and DESIGN.md).

    python tools/parity_report.py [C1 C2 C4 C5 large]  ->  one JSON line per case: first trace record that differs, max relative cost
    deviation, max absolute state deviation, flag mismatches.
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from pl_slam_plucker_b200 import abi, scene, solver  # noqa: E402
from oracle import loader as orc  # noqa: E402
import os  # noqa: E402


def report(name, s, P, opt, **extra):
    t0 = time.time(); r = s.solve(P, opt); tg = time.time() - t0
    t0 = time.time(); o = orc.solve(P, opt); to = time.time() - t0
    n = min(len(r.trace), len(o.trace))
    dec = [k for k in range(n) if any(r.trace[f][k] != o.trace[f][k] for f in ("stage", "iter", "trial", "accepted", "stop"))]
    m = dec[0] if dec else n
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if len(a) else 0.0
    fin = np.isfinite(o.trace["chi"][:m]) & np.isfinite(o.trace["chi_new"][:m])
    d = {"case": name, "n_free": int(P.n_free), "obs": int(P.n_obs), "trials_gpu": int(len(r.trace)), "trials_oracle": int(len(o.trace)), "first_decision_diff": (m if dec else None),
         "rho_at_diff": (float(o.trace["rho"][m]) if dec else None),
         "cost_rel": rel(r.trace["chi"][:m][fin], o.trace["chi"][:m][fin]), "cost_new_rel": rel(r.trace["chi_new"][:m][fin], o.trace["chi_new"][:m][fin]),
         "lambda_rel": rel(r.trace["lambda"][:m], o.trace["lambda"][:m]),
         "per_trial_cost_new_rel": [float(abs(r.trace["chi_new"][k] - o.trace["chi_new"][k]) / max(abs(o.trace["chi_new"][k]), 1e-300)) for k in range(m)],
         "pose_abs": float(np.abs(r.kf_T_wc - o.kf_T_wc).max()), "pt_abs": float(np.abs(r.pt_xyz - o.pt_xyz).max()),
         "ls_abs": float(np.abs(r.ls_orth - o.ls_orth).max()) if opt.profile != abi.PROFILE_H_END and P.n_ls else 0.0,
         "flag_mismatch": int((r.po_flags != o.po_flags).sum() + (r.lo_flags != o.lo_flags).sum()) if opt.profile == abi.PROFILE_G else 0,
         "chi2_rel": rel(r.po_chi2, np.maximum(o.po_chi2, 1e-12)) if opt.profile == abi.PROFILE_G else 0.0,
         "kernel_path": s.kernel_path(), "gpu_s": tg, "oracle_s": to}
    d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    what = sys.argv[1:] or ["C1", "C2", "C4", "large"]
    orc.set_threads(os.cpu_count() or 1)
    s = solver.LBASolver(0)
    for w in what:
        if w in ("C1", "C2", "C4"):
            P = scene.make_scene(int(w[1]))
            for q in (0, 1):
                report("%s G q%d" % (w, q), s, P, abi.Options(abi.PROFILE_G, q))
        elif w == "C5":
            P = scene.make_scene(5)
            for q in (0, 1):
                report("C5 G q%d (3+2 outer iterations)" % q, s, P, abi.Options(abi.PROFILE_G, q, iters_stage1=3, iters_stage2=2))
        elif w == "large":
            for i in range(12):
                rng = np.random.default_rng(5000 + i)
                nf = int(rng.integers(25, 72))
                kw = dict(n_kf_free=nf, n_kf_fixed=int(rng.integers(1, 3)), n_pt=int(rng.integers(25 * nf, 40 * nf)), n_ls=int(rng.integers(0, 8 * nf)),
                          mean_track=float(rng.uniform(3.5, 9.0)), seed=int(rng.integers(1, 10 ** 6)))
                if rng.random() < 0.25:
                    kw["loop_every"] = int(rng.integers(18, 24))
                P = scene.make_scene(1, **kw)
                report("large %d G faithful" % i, s, P, abi.Options(abi.PROFILE_G, 0), shape=kw)
    s.close()


if __name__ == "__main__":
    main()
