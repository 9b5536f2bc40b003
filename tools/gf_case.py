"""One draw of the G-faithful large-window sweep on every large-window solver against the oracle: python tools/gf_case.py <offset> <i>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pl_slam_plucker_b200 import abi, scene, solver
from oracle import loader as orc
off, i = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(5000 + i + off)
nf = int(rng.integers(25, 72))
kw = dict(n_kf_free=nf, n_kf_fixed=int(rng.integers(1, 3)), n_pt=int(rng.integers(25 * nf, 40 * nf)), n_ls=int(rng.integers(0, 8 * nf)),
          mean_track=float(rng.uniform(3.5, 9.0)), seed=int(rng.integers(1, 10 ** 6)))
if rng.random() < 0.25: kw["loop_every"] = int(rng.integers(18, 24))
print(kw)
P = scene.make_scene(1, **kw); opt = abi.Options(abi.PROFILE_G, 0)
o = orc.solve(P, opt)
res = {}
for name, env in (("default", {}), ("dense", {"PLBA_FORCE_DENSE": "1"}), ("band", {"PLBA_LARGE_SOLVER": "band"})):
    os.environ.update(env)
    s = solver.LBASolver(0)
    for k in env: os.environ.pop(k)
    r = s.solve(P, opt); res[name] = r
    n = min(len(r.trace), len(o.trace))
    dev = np.abs(r.trace["chi"][:n] / o.trace["chi"][:n] - 1)
    first = int(np.argmax(dev > 1e-9)) if (dev > 1e-9).any() else -1
    print("%-8s solver %-24s cost dev max %.2e (first trial above 1e-9: %d), state dev pts %.2e poses %.2e, trials %d / %d" % (
        name, s.kernel_path()["solver"], float(np.nanmax(dev)), first, float(np.abs(r.pt_xyz - o.pt_xyz).max()), float(np.abs(r.kf_T_wc - o.kf_T_wc).max()), len(r.trace), len(o.trace)))
    s.close()
a, b = res["default"], res["dense"]
n = min(len(a.trace), len(b.trace))
print("default vs dense: cost dev max %.2e, state dev %.2e" % (float(np.nanmax(np.abs(a.trace["chi"][:n] / b.trace["chi"][:n] - 1))), float(np.abs(a.pt_xyz - b.pt_xyz).max())))
# lambda / rho around the first deviating trial
dev = np.abs(a.trace["chi"][:n] / o.trace["chi"][:n] - 1)
k = int(np.argmax(dev > 1e-9)) if (dev > 1e-9).any() else 0
for t in range(max(0, k - 2), min(n, k + 2)):
    print("  trial", t, "oracle chi %.9g lambda %.3e rho %.3e acc %d | gpu chi %.9g lambda %.3e acc %d" % (o.trace["chi"][t], o.trace["lambda"][t], o.trace["rho"][t], o.trace["accepted"][t], a.trace["chi"][t], a.trace["lambda"][t], a.trace["accepted"][t]))
