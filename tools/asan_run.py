"""AddressSanitizer pass over the host emulation of every kernel family (tools/asan_emu.sh builds and runs it; compute-sanitizer is closed on the GPU pool)."""
import sys, os, ctypes as C
sys.path.insert(0,'/root/repo')
from pl_slam_plucker_b200 import abi, scene, solver, _lib
L=_lib.declare(C.CDLL(os.environ.get('PLBA_ASAN_SO', '/tmp/libplba_emu_asan.so')))
s=solver.LBASolver(0, lib=L)
for path in (1,2):
    s.set_kernel_path(path)
    for prof, lm in ((abi.PROFILE_G,0),(abi.PROFILE_H_END,1),(abi.PROFILE_H_PLK,0)):
        P=scene.make_scene(1, lib=L, n_kf_free=6, n_kf_fixed=2, n_pt=150, n_ls=40, line_mode=lm, seed=3)
        r=s.solve(P, abi.Options(prof,0)); print("path",path,"profile",prof,"trials",r.n_trials, flush=True)
    P=scene.make_scene(1, lib=L, n_kf_free=40, n_kf_fixed=2, n_pt=700, n_ls=150, seed=4)
    r=s.solve(P, abi.Options(abi.PROFILE_G,1, iters_stage1=2, iters_stage2=1)); print("bcr", s.kernel_path(), r.n_trials, flush=True)
    P2=scene.make_scene(1, lib=L, n_kf_free=33, n_kf_fixed=2, n_pt=500, n_ls=100, seed=5)     # last node smaller than the others
    r=s.solve(P2, abi.Options(abi.PROFILE_H_PLK,1)); print("bcr ragged", s.kernel_path(), r.n_trials, flush=True)
    Pl=scene.make_scene(1, lib=L, n_kf_free=22, n_kf_fixed=14, n_pt=300, n_ls=60, mean_track=34.0, seed=29)
    r=s.solve(Pl, abi.Options(abi.PROFILE_G,1, iters_stage1=2, iters_stage2=1)); print("long tracks", s.kernel_path(), r.n_trials, flush=True)
s.set_kernel_path(0)
os.environ["X"]="1"
L.plba_set_force_dense(s.h, 1)
r=s.solve(P, abi.Options(abi.PROFILE_G,1, iters_stage1=1, iters_stage2=1)); print("dense", s.kernel_path(), r.n_trials, flush=True)
L.plba_set_force_dense(s.h, 0)
probs=scene.make_batch(4,3,lib=L,n_pt=200,n_ls=50)
rc,rs=s.solve_batch(probs, abi.Options(abi.PROFILE_G,0)); print("batch",rc,[x.n_trials for x in rs], flush=True)
s.close(); print("done")
