"""Frame-to-frame pose tracking in Plücker mode through the C ABI (plba_track_solve): the ctypes face of
StereoFrameHandler::gaussNewtonOptimizationforPluker (src2/stereoFrameHandler.cpp:803-853) plus a synthetic stereo-frame
generator for tests and benchmarks.  CUDA only: the call goes to libplba.so and fails when it is missing."""
import ctypes as C

import numpy as np

_pd, _pb = C.POINTER(C.c_double), C.POINTER(C.c_uint8)


class plba_track_frame(C.Structure):
    _fields_ = [("n_pt", C.c_int32), ("n_ls", C.c_int32), ("DT", C.c_double * 12),
                ("pt_P", _pd), ("pt_obs", _pd), ("pt_inlier", _pb),
                ("ls_sP", _pd), ("ls_eP", _pd), ("ls_NDc", _pd), ("ls_obs", _pd), ("ls_seg", _pd), ("ls_sigma2", _pd), ("ls_inlier", _pb)]


class plba_track_options(C.Structure):
    _fields_ = [("cam", C.c_double * 4), ("homog_th", C.c_double), ("min_error", C.c_double), ("min_error_change", C.c_double),
                ("max_iters", C.c_int32), ("reserved", C.c_int32)]


class plba_track_result(C.Structure):
    _fields_ = [("DT", C.c_double * 12), ("DT_cov", C.c_double * 36), ("err", C.c_double), ("iters", C.c_int32), ("good", C.c_int32)]


def declare(L):
    L.plba_track_default_options.argtypes = [C.POINTER(plba_track_options)]
    L.plba_track_default_options.restype = None
    L.plba_track_solve.argtypes = [C.c_void_p, C.c_int32, C.POINTER(plba_track_frame), C.POINTER(plba_track_options), C.POINTER(plba_track_result)]
    L.plba_track_solve.restype = C.c_int
    return L


class Frame:
    """One matched stereo-frame pair: the arrays of plba_track_frame (include/plba.h), kept alive on the Python side."""

    def __init__(self, DT0, pt_P, pt_obs, ls_sP, ls_eP, ls_NDc, ls_obs, ls_seg, pt_inlier=None, ls_inlier=None, ls_sigma2=None):
        f8 = lambda a, w: np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1, w))
        self.DT0 = np.ascontiguousarray(np.asarray(DT0, np.float64).reshape(-1)[:12])
        self.pt_P, self.pt_obs = f8(pt_P, 3), f8(pt_obs, 2)
        self.ls_sP, self.ls_eP, self.ls_NDc, self.ls_obs, self.ls_seg = f8(ls_sP, 3), f8(ls_eP, 3), f8(ls_NDc, 6), f8(ls_obs, 4), f8(ls_seg, 4)
        self.pt_inlier = None if pt_inlier is None else np.ascontiguousarray(pt_inlier, np.uint8)
        self.ls_inlier = None if ls_inlier is None else np.ascontiguousarray(ls_inlier, np.uint8)
        self.ls_sigma2 = None if ls_sigma2 is None else np.ascontiguousarray(ls_sigma2, np.float64)

    def as_c(self):
        f = plba_track_frame()
        f.n_pt, f.n_ls = self.pt_P.shape[0], self.ls_NDc.shape[0]
        for i in range(12):
            f.DT[i] = float(self.DT0[i])
        for name, typ in (("pt_P", _pd), ("pt_obs", _pd), ("pt_inlier", _pb), ("ls_sP", _pd), ("ls_eP", _pd), ("ls_NDc", _pd), ("ls_obs", _pd),
                          ("ls_seg", _pd), ("ls_sigma2", _pd), ("ls_inlier", _pb)):
            a = getattr(self, name)
            setattr(f, name, C.cast(None, typ) if a is None or a.size == 0 else a.ctypes.data_as(typ))
        return f


class Options:
    def __init__(self, cam, max_iters=5, homog_th=1e-7, min_error=1e-7, min_error_change=1e-7):
        o = plba_track_options()
        for i in range(4):
            o.cam[i] = float(cam[i])
        o.homog_th, o.min_error, o.min_error_change, o.max_iters = homog_th, min_error, min_error_change, int(max_iters)
        self.c = o


def _unpack(res):
    return [{"DT": np.array(r.DT[:]).reshape(3, 4), "DT_cov": np.array(r.DT_cov[:]).reshape(6, 6), "err": float(r.err), "iters": int(r.iters), "good": int(r.good)} for r in res]


class Batch:
    """The C array of plba_track_frame for a list of frames, marshalled once (a caller that tracks many frames per call keeps it)."""

    def __init__(self, frames):
        self.frames = list(frames)
        self.n = len(self.frames)
        self.arr = (plba_track_frame * self.n)(*[f.as_c() for f in self.frames])
        self.res = (plba_track_result * self.n)()


def solve(solver, frames, opt, unpack=True):
    """Tracks a batch of independent frames (a list of Frame or a Batch) on the GPU of `solver` (an LBASolver): list of result dicts."""
    declare(solver.L)
    b = frames if isinstance(frames, Batch) else Batch(frames)
    solver._check(solver.L.plba_track_solve(solver.h, b.n, b.arr, C.byref(opt.c), b.res))
    return _unpack(b.res) if unpack else b.res


# ---- synthetic frames (tests / benchmarks) ---------------------------------------------------------------------------
def _exp_se3(x):
    """expmap_se3 (src2/auxiliar.cpp:124-141), translation first."""
    rho, w = np.asarray(x[:3], float), np.asarray(x[3:], float)
    th = np.linalg.norm(w)
    T = np.eye(4)
    if th < 1e-6:
        T[:3, 3] = rho
        return T
    s = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    T[:3, :3] = np.eye(3) + s * np.sin(th) + s @ s * (1 - np.cos(th))
    V = np.eye(3) + s * (1 - np.cos(th)) / th + s @ s * (th - np.sin(th)) / th
    T[:3, 3] = V @ rho
    return T


def make_frame(seed, n_pt=200, n_ls=80, cam=(435.2, 435.2, 367.2, 252.2), size=(752, 480), noise=0.5, outliers=0.05, motion=(0.05, 0.02), guess_noise=0.0):
    """A previous / current frame pair: 3-D points and segments seen in the previous camera frame, their pixels in the current one.
    Returns (Frame, DT_true 4x4).  DT maps previous-frame coordinates to current-frame coordinates."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = cam
    W, Hh = size

    def backproject(n):
        u, v, z = rng.uniform(20, W - 20, n), rng.uniform(20, Hh - 20, n), rng.uniform(2.0, 12.0, n)
        return np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1), np.stack([u, v], 1)

    def project(P):
        return np.stack([cx + fx * P[:, 0] / P[:, 2], cy + fy * P[:, 1] / P[:, 2]], 1)

    DT = _exp_se3(np.r_[rng.normal(size=3) * motion[0], rng.normal(size=3) * motion[1]])
    R, t = DT[:3, :3], DT[:3, 3]
    P, _ = backproject(n_pt)
    obs = project(P @ R.T + t) + rng.normal(size=(n_pt, 2)) * noise
    bad = rng.random(n_pt) < outliers
    obs[bad] += rng.uniform(-30, 30, (int(bad.sum()), 2))
    sP, spl = backproject(n_ls)
    d = rng.normal(size=(n_ls, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    eP = sP + d * rng.uniform(0.5, 2.0, (n_ls, 1))
    eP[:, 2] = np.maximum(eP[:, 2], 1.0)
    dd = (eP - sP) / np.linalg.norm(eP - sP, axis=1, keepdims=True)
    NDc = np.concatenate([np.cross(sP, dd), dd], 1)            # Plücker [n; d] in the previous frame, |d| = 1 (src/mapHandler.cpp:453-459)
    epl = project(eP)
    so = project(sP @ R.T + t) + rng.normal(size=(n_ls, 2)) * noise
    eo = project(eP @ R.T + t) + rng.normal(size=(n_ls, 2)) * noise
    badl = rng.random(n_ls) < outliers
    so[badl] += rng.uniform(-30, 30, (int(badl.sum()), 2))
    DT0 = np.eye(4) if guess_noise == 0.0 else _exp_se3(rng.normal(size=6) * guess_noise)
    return Frame(DT0[:3, :], P, obs, sP, eP, NDc, np.concatenate([so, eo], 1), np.concatenate([spl, epl], 1)), DT


# ---- creation of Plücker line landmarks (plba_create_lines) ---------------------------------------
_pi = C.POINTER(C.c_int32)


class plba_newline_batch(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_kf", C.c_int32), ("cam", C.c_double * 5), ("seg_l", _pd), ("seg_r", _pd), ("seg_curr", _pd),
                ("kf_prev", _pi), ("kf_curr", _pi), ("kf_T_wc", _pd)]


def _newline_batch(cam5, seg_l, seg_r, seg_curr, kf_prev, kf_curr, kf_T_wc):
    keep = [np.ascontiguousarray(seg_l, np.float64).reshape(-1, 4), np.ascontiguousarray(seg_r, np.float64).reshape(-1, 4),
            np.ascontiguousarray(seg_curr, np.float64).reshape(-1, 4), np.ascontiguousarray(kf_prev, np.int32), np.ascontiguousarray(kf_curr, np.int32),
            np.ascontiguousarray(kf_T_wc, np.float64).reshape(-1, 12)]
    B = plba_newline_batch()
    B.n, B.n_kf = keep[0].shape[0], keep[5].shape[0]
    for i in range(5):
        B.cam[i] = float(cam5[i])
    B.seg_l, B.seg_r, B.seg_curr = (a.ctypes.data_as(_pd) for a in keep[:3])
    B.kf_prev, B.kf_curr, B.kf_T_wc = keep[3].ctypes.data_as(_pi), keep[4].ctypes.data_as(_pi), keep[5].ctypes.data_as(_pd)
    return B, keep


def _newline_call(fn, first_args, cam5, seg_l, seg_r, seg_curr, kf_prev, kf_curr, kf_T_wc):
    B, keep = _newline_batch(cam5, seg_l, seg_r, seg_curr, kf_prev, kf_curr, kf_T_wc)
    n = B.n
    out = {"NDc": np.zeros((n, 6)), "NDw": np.zeros((n, 6)), "err_first": np.zeros(n), "err_curr": np.zeros(n), "accept": np.zeros(n, np.uint8)}
    rc = fn(*first_args, C.byref(B), out["NDc"].ctypes.data_as(_pd), out["NDw"].ctypes.data_as(_pd), out["err_first"].ctypes.data_as(_pd),
            out["err_curr"].ctypes.data_as(_pd), out["accept"].ctypes.data_as(_pb))
    return rc, out


def create_lines(solver, cam5, seg_l, seg_r, seg_curr, kf_prev, kf_curr, kf_T_wc):
    """Stereo-triangulated Plücker lines in the creating keyframe (NDc) and in the world (NDw), re-projection errors and the sqrt(5.991) gate."""
    L = solver.L
    L.plba_create_lines.argtypes = [C.c_void_p, C.POINTER(plba_newline_batch), _pd, _pd, _pd, _pd, _pb]
    L.plba_create_lines.restype = C.c_int
    rc, out = _newline_call(L.plba_create_lines, (solver.h,), cam5, seg_l, seg_r, seg_curr, kf_prev, kf_curr, kf_T_wc)
    solver._check(rc)
    return out


def make_line_candidates(seed, n=500, n_kf=6, cam5=(435.2, 435.2, 367.2, 252.2, 0.110078), noise=0.3, outliers=0.1):
    """3-D segments seen by a rectified stereo pair at keyframe `kf_prev` and by the left camera of `kf_curr`.
    Returns the arguments of create_lines plus the true world Plücker lines (unit direction)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, b = cam5
    T = np.zeros((n_kf, 3, 4))
    for k in range(n_kf):
        T[k] = _exp_se3(np.r_[0.3 * k + rng.normal() * 0.02, rng.normal(size=2) * 0.02, rng.normal(size=3) * 0.02])[:3]      # camera -> world
    kp = rng.integers(0, n_kf - 1, n).astype(np.int32); kc = (kp + 1).astype(np.int32)
    u, v, z = rng.uniform(60, 700, n), rng.uniform(40, 440, n), rng.uniform(2.0, 10.0, n)
    sP = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)                    # in the creating camera frame
    d = rng.normal(size=(n, 3)); d[:, 2] *= 0.3; d /= np.linalg.norm(d, axis=1, keepdims=True)
    eP = sP + d * rng.uniform(0.4, 1.5, (n, 1)); eP[:, 2] = np.maximum(eP[:, 2], 1.0)
    proj = lambda P, off=0.0: np.stack([cx + fx * (P[:, 0] - off) / P[:, 2], cy + fy * P[:, 1] / P[:, 2]], 1)
    seg_l = np.concatenate([proj(sP), proj(eP)], 1) + rng.normal(size=(n, 4)) * noise
    seg_r = np.concatenate([proj(sP, b), proj(eP, b)], 1) + rng.normal(size=(n, 4)) * noise
    seg_r[:, 1], seg_r[:, 3] = seg_l[:, 1], seg_l[:, 3]                              # same rows after the interpolation of :367-368
    Rp, tp = T[kp][:, :, :3], T[kp][:, :, 3]
    sW = np.einsum("nij,nj->ni", Rp, sP) + tp; eW = np.einsum("nij,nj->ni", Rp, eP) + tp
    Rc, tc = T[kc][:, :, :3], T[kc][:, :, 3]
    toc = lambda Pw: np.einsum("nji,nj->ni", Rc, Pw - tc)
    seg_curr = np.concatenate([proj(toc(sW)), proj(toc(eW))], 1) + rng.normal(size=(n, 4)) * noise
    bad = rng.random(n) < outliers
    seg_curr[bad] += rng.uniform(-25, 25, (int(bad.sum()), 4))
    dW = (eW - sW) / np.linalg.norm(eW - sW, axis=1, keepdims=True)
    truth = np.concatenate([np.cross(sW, dW), dW], 1)
    return (cam5, seg_l, seg_r, seg_curr, kp, kc, T.reshape(n_kf, 12)), truth
