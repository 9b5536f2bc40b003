import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pl_slam_plucker_b200 import abi, scene, solver
P = scene.make_scene(5); s = solver.LBASolver(0)
opt = abi.Options(abi.PROFILE_G, 1, iters_stage1=3, iters_stage2=2)
r = s.solve(P, opt)
tr = r.trace
print("trace", [(int(t["stage"]), int(t["iter"]), int(t["accepted"]), float(t["chi"]), float(t["chi_new"]), float(t["rho"]), float(t["lambda"])) for t in tr])
print("median chi2 pt/ls", float(np.median(r.po_chi2)), float(np.median(r.lo_chi2)), "mean", float(np.mean(r.po_chi2)), "bad frac", float(((r.po_flags & 2) != 0).mean()), float(((r.lo_flags & 2) != 0).mean()))
print("lo flags vs chi2 mismatch", int((((r.lo_flags & abi.OBS_BAD) != 0) != (r.lo_chi2 > opt.chi2_gate)).sum()), "finite", bool(np.isfinite(r.pt_xyz).all()), bool(np.isfinite(r.kf_T_wc).all()))
print("timing", s.timing())
