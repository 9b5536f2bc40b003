"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracle (oracle/libplba_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from pl_slam_plucker_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libplba_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("plba_oracle.cpp", "track_oracle.cpp", "refmath.h", "smallmat.h")] + [os.path.join(_HERE, "..", "include", "plba.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libplba_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        pd = C.POINTER(C.c_double)
        L.plba_oracle_solve.argtypes = [C.POINTER(abi.plba_problem), C.POINTER(abi.plba_options), C.POINTER(abi.plba_result), C.c_int]
        L.plba_oracle_solve.restype = C.c_int
        L.plba_oracle_solve_batch.argtypes = [C.c_int, C.POINTER(abi.plba_problem), C.POINTER(abi.plba_options), C.POINTER(abi.plba_result), C.c_int]
        L.plba_oracle_solve_batch.restype = C.c_int
        L.plba_oracle_reduced_system_G.argtypes = [C.POINTER(abi.plba_problem), C.POINTER(abi.plba_options), C.c_double, pd, pd, pd]
        L.plba_oracle_reduced_system_G.restype = C.c_int
        L.plba_oracle_set_threads.argtypes = [C.c_int]
        L.plba_oracle_get_threads.restype = C.c_int
        _LIB = L
    return _LIB


def set_threads(n):
    lib().plba_oracle_set_threads(int(n))
    return lib().plba_oracle_get_threads()


def solve(prob, opt, dense=False, trace_cap=256):
    """CPU oracle LBA on one window.  dense=True: literal full-system solve (small problems)."""
    res = abi.Result(prob, trace_cap)
    pc = prob.as_c()
    rc = lib().plba_oracle_solve(C.byref(pc), C.byref(opt.c), C.byref(res.c), 1 if dense else 0)
    res.rc = rc
    return res


def solve_batch(probs, opt, trace_cap=64):
    n = len(probs)
    arr = (abi.plba_problem * n)(*[p.as_c() for p in probs])
    results = [abi.Result(p, trace_cap) for p in probs]
    rarr = (abi.plba_result * n)(*[r.c for r in results])
    rc = lib().plba_oracle_solve_batch(n, arr, C.byref(opt.c), rarr, 0)
    for i, r in enumerate(results):
        r.c = rarr[i]
        r.rc = rarr[i].status
    return rc, results


def reduced_system_G(prob, opt, lam):
    nf = prob.n_free
    S = np.zeros((nf, nf, 6, 6)); g = np.zeros(6 * nf); chi = C.c_double(0)
    pc = prob.as_c()
    pd = C.POINTER(C.c_double)
    rc = lib().plba_oracle_reduced_system_G(C.byref(pc), C.byref(opt.c), float(lam), S.ctypes.data_as(pd), g.ctypes.data_as(pd), C.byref(chi))
    assert rc == 0
    return S, g, chi.value


def _v(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def expmap_se3(x):
    x = _v(x); T = np.zeros(16); lib().plba_oracle_expmap_se3(_p(x), _p(T)); return T.reshape(4, 4)


def logmap_se3(T):
    T = _v(T).reshape(16); x = np.zeros(6); lib().plba_oracle_logmap_se3(_p(T), _p(x)); return x


def inverse_se3(T):
    T = _v(T).reshape(16); o = np.zeros(16); lib().plba_oracle_inverse_se3(_p(T), _p(o)); return o.reshape(4, 4)


def pluker_to_orth(pl):
    pl = _v(pl); o = np.zeros(4); lib().plba_oracle_pluker_to_orth(_p(pl), _p(o)); return o


def orth_to_pluker(o):
    o = _v(o); pl = np.zeros(6); lib().plba_oracle_orth_to_pluker(_p(o), _p(pl)); return pl


def orth_UW_jac(pl, q7=False):
    pl = _v(pl); U = np.zeros(9); W = np.zeros(4); J = np.zeros(24)
    lib().plba_oracle_orth_UW_jac(_p(pl), C.c_int(1 if q7 else 0), _p(U), _p(W), _p(J)); return U.reshape(3, 3), W.reshape(2, 2), J.reshape(6, 4)


def transform_pluker(T, pl):
    T = _v(T).reshape(16); pl = _v(pl); o = np.zeros(6); lib().plba_oracle_transform_pluker(_p(T), _p(pl), _p(o)); return o


def update_orth(D, d):
    D = _v(D); d = _v(d); o = np.zeros(4); lib().plba_oracle_update_orth(_p(D), _p(d), _p(o)); return o


def pose_oplus(T, d):
    T = _v(T).reshape(16); d = _v(d); o = np.zeros(16); lib().plba_oracle_pose_oplus(_p(T), _p(d), _p(o)); return o.reshape(4, 4)


def point_edge(cam, Tcw, Pw, obs):
    cam = _v(cam); T = _v(Tcw).reshape(16); Pw = _v(Pw); obs = _v(obs)
    e = np.zeros(2); Ji = np.zeros(6); Jj = np.zeros(12)
    lib().plba_oracle_point_edge(_p(cam), _p(T), _p(Pw), _p(obs), _p(e), _p(Ji), _p(Jj))
    return e, Ji.reshape(2, 3), Jj.reshape(2, 6)


def line_edge(cam, Tcw, orth, obs, faithful=True):
    cam = _v(cam); T = _v(Tcw).reshape(16); orth = _v(orth); obs = _v(obs)
    e = np.zeros(2); Ji = np.zeros(8); Jj = np.zeros(12)
    lib().plba_oracle_line_edge(_p(cam), _p(T), _p(orth), _p(obs), C.c_int(1 if faithful else 0), _p(e), _p(Ji), _p(Jj))
    return e, Ji.reshape(2, 4), Jj.reshape(2, 6)


def h_term(kind, cam, Tiw, lm, obs, th=1e-7, fixed=False):
    cam = _v(cam); T = _v(Tiw).reshape(16); lm6 = np.zeros(6); lm6[:len(lm)] = lm; ob4 = np.zeros(4); ob4[:len(obs)] = obs
    Jp = np.zeros(6); Jl = np.zeros(6); rw = np.zeros(2)
    lib().plba_oracle_h_term(C.c_int(kind), _p(cam), _p(T), _p(lm6), _p(ob4), C.c_double(th), C.c_int(1 if fixed else 0), _p(Jp), _p(Jl), _p(rw))
    return Jp, Jl, rw[0], rw[1]


def track_solve(frame, opt):
    """CPU oracle of the Plücker-mode pose tracking (oracle/track_oracle.cpp) on one frame: result dict as tracking.solve."""
    from pl_slam_plucker_b200 import tracking as trk
    L = lib()
    L.plba_track_oracle.argtypes = [C.POINTER(trk.plba_track_frame), C.POINTER(trk.plba_track_options), C.POINTER(trk.plba_track_result)]
    L.plba_track_oracle.restype = C.c_int
    fc = frame.as_c(); res = trk.plba_track_result()
    L.plba_track_oracle(C.byref(fc), C.byref(opt.c), C.byref(res))
    return trk._unpack([res])[0]


def create_lines(cam5, seg_l, seg_r, seg_curr, kf_prev, kf_curr, kf_T_wc):
    """CPU oracle of the Plücker line-landmark creation (oracle/track_oracle.cpp): dict as tracking.create_lines."""
    from pl_slam_plucker_b200 import tracking as trk
    L = lib()
    pd, pb = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    L.plba_create_lines_oracle.argtypes = [C.POINTER(trk.plba_newline_batch), pd, pd, pd, pd, pb]
    L.plba_create_lines_oracle.restype = C.c_int
    rc, out = trk._newline_call(L.plba_create_lines_oracle, (), cam5, seg_l, seg_r, seg_curr, kf_prev, kf_curr, kf_T_wc)
    return out
