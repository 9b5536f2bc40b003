"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/_ref/libref_g2o_types.so = the REFERENCE's own g2o_types/g2o_types.h
and src/mapFeatures.cpp compiled unmodified against stand-in Eigen / g2o / OpenCV headers (oracle/ref_shim/, oracle/ref_g2o_types.cpp,
oracle/ref_map_features.cpp).

It can only be BUILT where /root/reference exists (this container); the built library travels to the GPU box with the
snapshot.  Only tests/ (and tests/golden/make_ref_golden.py) import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_g2o_types.so")
REF_HEADER = "/root/reference/g2o_types/g2o_types.h"
_LIB = None


def available():
    return os.path.exists(SO) or os.path.exists(REF_HEADER)


def build(force=False):
    """Compile the reference header from where it lies; a no-op (returning None) where /root/reference is absent."""
    if not os.path.exists(REF_HEADER):
        return SO if os.path.exists(SO) else None
    srcs = [REF_HEADER, "/root/reference/src/mapFeatures.cpp", os.path.join(_HERE, "ref_g2o_types.cpp"), os.path.join(_HERE, "ref_map_features.cpp"),
            os.path.join(_HERE, "ref_shim", "standin.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return SO


def lib():
    global _LIB
    if _LIB is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libref_g2o_types.so is absent and /root/reference is not here to build it")
        _LIB = C.CDLL(SO)
        _LIB.ref_point_edge.restype = C.c_int
    return _LIB


def _v(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pose_oplus(T, d):
    T = _v(T).reshape(16); d = _v(d); o = np.zeros(16); lib().ref_pose_oplus(_p(T), _p(d), _p(o)); return o.reshape(4, 4)


def line_oplus(D, d):
    D = _v(D); d = _v(d); o = np.zeros(4); lib().ref_line_oplus(_p(D), _p(d), _p(o)); return o


def point_oplus(P, d):
    P = _v(P); d = _v(d); o = np.zeros(3); lib().ref_point_oplus(_p(P), _p(d), _p(o)); return o


def point_edge(cam, Tcw, Pw, obs):
    cam = _v(cam); T = _v(Tcw).reshape(16); Pw = _v(Pw); obs = _v(obs)
    e = np.zeros(2); Ji = np.zeros(6); Jj = np.zeros(12)
    pos = lib().ref_point_edge(_p(cam), _p(T), _p(Pw), _p(obs), _p(e), _p(Ji), _p(Jj))
    return e, Ji.reshape(2, 3), Jj.reshape(2, 6), bool(pos)


def line_edge(cam, Tcw, orth, obs):
    cam = _v(cam); T = _v(Tcw).reshape(16); orth = _v(orth); obs = _v(obs)
    e = np.zeros(4); Ji = np.zeros(16); Jj = np.zeros(24); chi2 = C.c_double(0)
    lib().ref_line_edge(_p(cam), _p(T), _p(orth), _p(obs), _p(e), _p(Ji), _p(Jj), C.byref(chi2))
    return e, Ji.reshape(4, 4), Jj.reshape(4, 6), chi2.value


def orth_to_pluker(o):
    o = _v(o); pl = np.zeros(6); lib().ref_orth_to_pluker(_p(o), _p(pl)); return pl


def orth_UW_jac(plk):
    plk = _v(plk); U = np.zeros(9); W = np.zeros(4); J = np.zeros(24)
    lib().ref_orth_UW_jac(_p(plk), _p(U), _p(W), _p(J)); return U.reshape(3, 3), W.reshape(2, 2), J.reshape(6, 4)


def transform_pluker(Tcw, plk):
    T = _v(Tcw).reshape(16); plk = _v(plk); o = np.zeros(6); lib().ref_transform_pluker(_p(T), _p(plk), _p(o)); return o


# ---- MapLine:: copies (src/mapFeatures.cpp:186-266) ----
def ml_pluker_to_orth(plk):
    plk = _v(plk); o = np.zeros(4); lib().ref_ml_pluker_to_orth(_p(plk), _p(o)); return o


def ml_orth_to_pluker(o):
    o = _v(o); pl = np.zeros(6); lib().ref_ml_orth_to_pluker(_p(o), _p(pl)); return pl


def ml_UW_jac(plk):
    plk = _v(plk); U = np.zeros(9); W = np.zeros(4); J = np.zeros(24)
    lib().ref_ml_UW_jac(_p(plk), _p(U), _p(W), _p(J)); return U.reshape(3, 3), W.reshape(2, 2), J.reshape(6, 4)
