"""Small cases of every kernel family for compute-sanitizer (memcheck): warp + chunk assembly, small / BCR / banded / dense solvers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver
s = solver.LBASolver(0)
for path in (1, 2):
    s.set_kernel_path(path)
    for prof, lm in ((abi.PROFILE_G, 0), (abi.PROFILE_H_END, 1), (abi.PROFILE_H_PLK, 0)):
        P = scene.make_scene(1, n_kf_free=6, n_kf_fixed=2, n_pt=150, n_ls=40, line_mode=lm, seed=3)
        r = s.solve(P, abi.Options(prof, 0)); print("path", path, "profile", prof, "trials", r.n_trials)
    P = scene.make_scene(1, n_kf_free=40, n_kf_fixed=2, n_pt=700, n_ls=150, seed=4)          # block cyclic reduction
    r = s.solve(P, abi.Options(abi.PROFILE_G, 1, iters_stage1=2, iters_stage2=1)); print("bcr", s.kernel_path(), r.n_trials)
s.set_kernel_path(0)
s.set_force_dense(True)
r = s.solve(P, abi.Options(abi.PROFILE_G, 1, iters_stage1=1, iters_stage2=1)); print("dense", s.kernel_path(), r.n_trials)
s.set_force_dense(False)
probs = scene.make_batch(4, 3, n_pt=200, n_ls=50)
rc, rs = s.solve_batch(probs, abi.Options(abi.PROFILE_G, 0)); print("batch", rc, [x.n_trials for x in rs])
s.close(); print("done")
