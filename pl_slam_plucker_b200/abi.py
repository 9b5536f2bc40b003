"""ctypes mirror of include/plba.h (data layouts only; no compute).

`Problem` / `Options` / `Result` are NumPy-backed views of plba_problem / plba_options / plba_result, the flattened
form of what MapHandler::localBundleAdjustment() gathers (reference src/mapHandler.cpp:1392-1502).
"""
import ctypes as C
import numpy as np

PROFILE_G, PROFILE_H_END, PROFILE_H_PLK = 0, 1, 2
QUIRKS_FAITHFUL, QUIRKS_FIXED = 0, 1
SHELL_LBA, SHELL_GBA = 0, 1
OK, DISCARDED, E_ARG, E_CUDA, E_NUMERIC, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
OBS_LEVEL1, OBS_BAD, OBS_NEGDEPTH = 1, 2, 4

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pb = C.POINTER(C.c_uint8)


class plba_problem(C.Structure):
    _fields_ = [("n_kf", C.c_int32), ("n_free", C.c_int32), ("n_pt", C.c_int32), ("n_ls", C.c_int32),
                ("n_pobs", C.c_int32), ("n_lobs", C.c_int32), ("cam", C.c_double * 4),
                ("kf_T_wc", _pd), ("kf_slot", _pi), ("x_pose", _pd), ("pt_xyz", _pd), ("ls_plk", _pd), ("ls_end", _pd),
                ("po_lm", _pi), ("po_kf", _pi), ("po_uv", _pd), ("po_sig2", _pd),
                ("lo_lm", _pi), ("lo_kf", _pi), ("lo_ab", _pd), ("lo_sig2", _pd)]


class plba_options(C.Structure):
    _fields_ = [("profile", C.c_int32), ("quirks", C.c_int32),
                ("lambda_lba_lm", C.c_double), ("lambda_lba_k", C.c_double), ("max_iters_lba", C.c_int32), ("shell", C.c_int32),
                ("homog_th", C.c_double), ("min_error", C.c_double), ("min_error_change", C.c_double),
                ("huber_delta", C.c_double), ("chi2_gate", C.c_double), ("iters_stage1", C.c_int32), ("iters_stage2", C.c_int32),
                ("lm_tau", C.c_double), ("lm_max_trials", C.c_int32), ("reserved1", C.c_int32)]


class plba_trace_rec(C.Structure):
    _fields_ = [("window", C.c_int32), ("stage", C.c_int32), ("iter", C.c_int32), ("trial", C.c_int32),
                ("accepted", C.c_int32), ("stop", C.c_int32),
                ("chi", C.c_double), ("chi_new", C.c_double), ("rho", C.c_double), ("lambda_", C.c_double),
                ("scale", C.c_double), ("dx_norm", C.c_double), ("err_pt", C.c_double), ("err_ls", C.c_double)]


TRACE_DTYPE = np.dtype([("window", "i4"), ("stage", "i4"), ("iter", "i4"), ("trial", "i4"), ("accepted", "i4"), ("stop", "i4"),
                        ("chi", "f8"), ("chi_new", "f8"), ("rho", "f8"), ("lambda", "f8"), ("scale", "f8"),
                        ("dx_norm", "f8"), ("err_pt", "f8"), ("err_ls", "f8")])
assert TRACE_DTYPE.itemsize == C.sizeof(plba_trace_rec)


class plba_result(C.Structure):
    _fields_ = [("kf_T_wc", _pd), ("x_pose", _pd), ("pt_xyz", _pd), ("ls_plk", _pd), ("ls_orth", _pd), ("ls_end", _pd),
                ("pt_inlier", _pb), ("ls_inlier", _pb), ("po_chi2", _pd), ("lo_chi2", _pd), ("po_flags", _pb), ("lo_flags", _pb),
                ("trace", C.POINTER(plba_trace_rec)), ("trace_cap", C.c_int32), ("n_trace", C.c_int32),
                ("status", C.c_int32), ("n_trials", C.c_int32)]


class plba_timing(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_assemble", C.c_double), ("ms_solve", C.c_double), ("ms_update", C.c_double),
                ("ms_other", C.c_double), ("n_launches", C.c_int64), ("n_assemble", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("n_launches_run", C.c_int64), ("n_assemble_run", C.c_int64), ("n_trials_run", C.c_int64),
                ("ms_host_prep", C.c_double), ("ms_host_unpack", C.c_double)]


class plba_scene_spec(C.Structure):
    _fields_ = [("n_kf_free", C.c_int32), ("n_kf_fixed", C.c_int32), ("n_pt", C.c_int32), ("n_ls", C.c_int32),
                ("mean_track", C.c_double), ("width", C.c_int32), ("height", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("kf_spacing", C.c_double), ("depth_min", C.c_double), ("depth_max", C.c_double),
                ("pixel_noise", C.c_double), ("outlier_frac", C.c_double),
                ("pose_rot_noise", C.c_double), ("pose_trans_noise", C.c_double), ("pt_noise", C.c_double), ("ls_noise", C.c_double),
                ("seed", C.c_uint64), ("line_mode", C.c_int32), ("loop_every", C.c_int32)]


def _f8(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _i4(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, typ):
    return C.cast(None, typ) if a is None or a.size == 0 and False else a.ctypes.data_as(typ)


class Problem:
    """One flattened LBA window (see plba_problem in include/plba.h)."""

    def __init__(self, cam, kf_T_wc, kf_slot, pt_xyz, po_lm, po_kf, po_uv, ls_plk=None, ls_end=None, lo_lm=None, lo_kf=None,
                 lo_ab=None, x_pose=None, po_sig2=None, lo_sig2=None):
        self.cam = _f8(cam, (4,))
        self.kf_T_wc = _f8(kf_T_wc, (-1, 12))
        self.kf_slot = _i4(kf_slot)
        self.n_kf = int(self.kf_T_wc.shape[0])
        self.n_free = int((self.kf_slot >= 0).sum())
        self.pt_xyz = _f8(pt_xyz if pt_xyz is not None else np.zeros((0, 3)), (-1, 3))
        self.ls_plk = _f8(ls_plk, (-1, 6))
        self.ls_end = _f8(ls_end, (-1, 6))
        self.n_pt = int(self.pt_xyz.shape[0])
        self.n_ls = int(self.ls_plk.shape[0]) if self.ls_plk is not None else (int(self.ls_end.shape[0]) if self.ls_end is not None else 0)
        self.po_lm = _i4(po_lm if po_lm is not None else np.zeros(0, np.int32))
        self.po_kf = _i4(po_kf if po_kf is not None else np.zeros(0, np.int32))
        self.po_uv = _f8(po_uv if po_uv is not None else np.zeros((0, 2)), (-1, 2))
        self.lo_lm = _i4(lo_lm if lo_lm is not None else np.zeros(0, np.int32))
        self.lo_kf = _i4(lo_kf if lo_kf is not None else np.zeros(0, np.int32))
        self.lo_ab = _f8(lo_ab if lo_ab is not None else np.zeros((0, 4)), (-1, 4))
        self.x_pose = _f8(x_pose, (-1, 6))
        self.po_sig2 = _f8(po_sig2)
        self.lo_sig2 = _f8(lo_sig2)
        self.n_pobs = int(self.po_lm.shape[0])
        self.n_lobs = int(self.lo_lm.shape[0])

    @property
    def n_obs(self):
        return self.n_pobs + self.n_lobs

    def as_c(self):
        p = plba_problem()
        p.n_kf, p.n_free, p.n_pt, p.n_ls, p.n_pobs, p.n_lobs = self.n_kf, self.n_free, self.n_pt, self.n_ls, self.n_pobs, self.n_lobs
        for i in range(4):
            p.cam[i] = float(self.cam[i])
        for name, typ in (("kf_T_wc", _pd), ("kf_slot", _pi), ("x_pose", _pd), ("pt_xyz", _pd), ("ls_plk", _pd), ("ls_end", _pd),
                          ("po_lm", _pi), ("po_kf", _pi), ("po_uv", _pd), ("po_sig2", _pd),
                          ("lo_lm", _pi), ("lo_kf", _pi), ("lo_ab", _pd), ("lo_sig2", _pd)):
            a = getattr(self, name)
            setattr(p, name, C.cast(None, typ) if a is None else a.ctypes.data_as(typ))
        return p

    def subset_landmarks(self, pt_mask, ls_mask):
        """Landmark shard (all KFs kept, landmarks re-indexed densely): the multi-GPU partition of SURVEY.md §8e."""
        pt_idx = np.flatnonzero(pt_mask); ls_idx = np.flatnonzero(ls_mask)
        pmap = -np.ones(self.n_pt, np.int64); pmap[pt_idx] = np.arange(pt_idx.size)
        lmap = -np.ones(max(self.n_ls, 1), np.int64); lmap[ls_idx] = np.arange(ls_idx.size)
        po = pt_mask[self.po_lm] if self.n_pobs else np.zeros(0, bool)
        lo = ls_mask[self.lo_lm] if self.n_lobs else np.zeros(0, bool)
        return Problem(self.cam, self.kf_T_wc, self.kf_slot, self.pt_xyz[pt_idx], pmap[self.po_lm[po]], self.po_kf[po], self.po_uv[po],
                       ls_plk=None if self.ls_plk is None else self.ls_plk[ls_idx],
                       ls_end=None if self.ls_end is None else self.ls_end[ls_idx],
                       lo_lm=lmap[self.lo_lm[lo]], lo_kf=self.lo_kf[lo], lo_ab=self.lo_ab[lo], x_pose=self.x_pose,
                       po_sig2=None if self.po_sig2 is None else self.po_sig2[po],
                       lo_sig2=None if self.lo_sig2 is None else self.lo_sig2[lo])


class Options:
    """LM schedule + robust weighting; defaults are the reference's (see plba_options)."""

    def __init__(self, profile=PROFILE_G, quirks=QUIRKS_FAITHFUL, **kw):
        o = plba_options()
        o.profile, o.quirks = profile, quirks
        o.lambda_lba_lm, o.lambda_lba_k, o.max_iters_lba = 1e-5, 10.0, 15       # src/slamConfig.cpp:65-67
        o.homog_th, o.min_error, o.min_error_change = 1e-7, 1e-7, 1e-7           # src2/config.cpp:80-85
        o.huber_delta = float(np.float32(np.sqrt(5.991)))                        # src/mapHandler.cpp:5978 (float, Q13)
        o.chi2_gate = 5.991
        o.iters_stage1, o.iters_stage2 = 5, 10                                   # :6122, :6152
        o.lm_tau, o.lm_max_trials = 1e-5, 10                                     # g2o defaults
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
        self.c = o

    def __getattr__(self, k):
        return getattr(self.__dict__["c"], k)


class Result:
    """Caller-owned output buffers for one window."""

    def __init__(self, prob, trace_cap=256):
        self.kf_T_wc = np.zeros((prob.n_kf, 12))
        self.x_pose = np.zeros((prob.n_free, 6))
        self.pt_xyz = np.zeros((prob.n_pt, 3))
        self.ls_plk = np.zeros((prob.n_ls, 6))
        self.ls_orth = np.zeros((prob.n_ls, 4))
        self.ls_end = np.zeros((prob.n_ls, 6))
        self.pt_inlier = np.zeros(prob.n_pt, np.uint8)
        self.ls_inlier = np.zeros(prob.n_ls, np.uint8)
        self.po_chi2 = np.zeros(prob.n_pobs)
        self.lo_chi2 = np.zeros(prob.n_lobs)
        self.po_flags = np.zeros(prob.n_pobs, np.uint8)
        self.lo_flags = np.zeros(prob.n_lobs, np.uint8)
        self._trace = np.zeros(trace_cap, TRACE_DTYPE)
        r = plba_result()
        for name in ("kf_T_wc", "x_pose", "pt_xyz", "ls_plk", "ls_orth", "ls_end", "po_chi2", "lo_chi2"):
            setattr(r, name, getattr(self, name).ctypes.data_as(_pd))
        for name in ("pt_inlier", "ls_inlier", "po_flags", "lo_flags"):
            setattr(r, name, getattr(self, name).ctypes.data_as(_pb))
        r.trace = self._trace.ctypes.data_as(C.POINTER(plba_trace_rec))
        r.trace_cap = trace_cap
        self.c = r

    @property
    def trace(self):
        return self._trace[:min(self.c.n_trace, self.c.trace_cap)]

    @property
    def status(self):
        return self.c.status

    @property
    def n_trials(self):
        return self.c.n_trials


def declare_common(lib):
    """argtypes for the entry points that exist in both the product library and (a subset) the oracle."""
    lib.plba_scene_preset.argtypes = [C.c_int32, C.POINTER(plba_scene_spec)]
    lib.plba_scene_preset.restype = None
    lib.plba_scene_create.argtypes = [C.POINTER(plba_scene_spec), C.POINTER(C.c_void_p)]
    lib.plba_scene_create.restype = C.c_int
    lib.plba_scene_problem.argtypes = [C.c_void_p, C.POINTER(plba_problem)]
    lib.plba_scene_problem.restype = C.c_int
    lib.plba_scene_truth.argtypes = [C.c_void_p, C.POINTER(_pd), C.POINTER(_pd), C.POINTER(_pd)]
    lib.plba_scene_truth.restype = C.c_int
    lib.plba_scene_destroy.argtypes = [C.c_void_p]
    lib.plba_scene_destroy.restype = None
