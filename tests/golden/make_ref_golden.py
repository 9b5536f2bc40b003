"""Generates tests/golden/ref_g2o_types.npz: outputs of the REFERENCE's own code (g2o_types/g2o_types.h and src/mapFeatures.cpp,
compiled unmodified into oracle/_ref/libref_g2o_types.so, see oracle/Makefile target `ref`) on seeded inputs.

Run in the build container (where /root/reference exists):  python tests/golden/make_ref_golden.py
The GPU box has no /root/reference; there the oracle is checked against these committed vectors (tests/test_ref_pin.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_loader as ref  # noqa: E402

N = 256
CAMS = np.array([[435.2, 435.2, 367.2, 252.2], [718.856, 718.856, 607.1928, 185.2157]])


def rand_T(rng):
    w = rng.normal(size=3) * 0.6
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    T[:3, 3] = rng.normal(size=3) * 2.0
    return T


def inputs(seed=20261018):
    rng = np.random.default_rng(seed)
    d = {}
    d["cam"] = CAMS[rng.integers(0, 2, N)]
    d["T"] = np.array([rand_T(rng) for _ in range(N)])
    d["Pw"] = rng.normal(size=(N, 3)) * 3 + np.array([0, 0, 9.0])
    d["Pw"][:8, 2] = -5.0                                        # behind the camera for T ~ I: isDepthPositive both ways
    d["uv"] = rng.uniform(0, 700, (N, 2))
    d["orth"] = rng.uniform(-1.4, 1.4, (N, 4)); d["orth"][:, 3] = rng.uniform(0.02, 1.55, N)
    d["ab"] = rng.uniform(0, 900, (N, 4))
    d["d6"] = rng.normal(size=(N, 6)) * 0.1
    d["d6"][0] = 0.0; d["d6"][1, 3:] = 1e-12; d["d6"][2, 3:] = [3e-11, 0, 0]; d["d6"][3, 3:] = 2.0      # Taylor branch (theta < 1e-10) and a large rotation
    d["d4"] = rng.normal(size=(N, 4)) * 0.1
    d["d4"][0] = 0.0; d["d4"][1] = [1.0, -1.2, 0.8, 0.5]
    d["d3"] = rng.normal(size=(N, 3))
    n = rng.normal(size=(N, 3)); dd = rng.normal(size=(N, 3)); dd -= n * (np.sum(n * dd, 1) / np.sum(n * n, 1))[:, None]
    d["plk"] = np.c_[n * rng.uniform(0.1, 10, (N, 1)), dd]
    return d


def outputs(d):
    o = {k: [] for k in ("pe_e", "pe_Ji", "pe_Jj", "pe_pos", "le_e", "le_Ji", "le_Jj", "le_chi2", "pose_oplus", "line_oplus", "point_oplus", "o2p", "U", "W", "J",
                         "tp", "ml_p2o", "ml_o2p", "ml_U", "ml_W", "ml_J")}
    for i in range(N):
        e, Ji, Jj, pos = ref.point_edge(d["cam"][i], d["T"][i], d["Pw"][i], d["uv"][i])
        o["pe_e"].append(e); o["pe_Ji"].append(Ji); o["pe_Jj"].append(Jj); o["pe_pos"].append(pos)
        e, Ji, Jj, chi = ref.line_edge(d["cam"][i], d["T"][i], d["orth"][i], d["ab"][i])
        o["le_e"].append(e); o["le_Ji"].append(Ji); o["le_Jj"].append(Jj); o["le_chi2"].append(chi)
        o["pose_oplus"].append(ref.pose_oplus(d["T"][i], d["d6"][i]))
        o["line_oplus"].append(ref.line_oplus(d["orth"][i], d["d4"][i]))
        o["point_oplus"].append(ref.point_oplus(d["Pw"][i], d["d3"][i]))
        o["o2p"].append(ref.orth_to_pluker(d["orth"][i]))
        U, W, J = ref.orth_UW_jac(d["plk"][i]); o["U"].append(U); o["W"].append(W); o["J"].append(J)
        o["tp"].append(ref.transform_pluker(d["T"][i], d["plk"][i]))
        o["ml_p2o"].append(ref.ml_pluker_to_orth(d["plk"][i])); o["ml_o2p"].append(ref.ml_orth_to_pluker(d["orth"][i]))
        U, W, J = ref.ml_UW_jac(d["plk"][i]); o["ml_U"].append(U); o["ml_W"].append(W); o["ml_J"].append(J)
    return {k: np.array(v) for k, v in o.items()}


if __name__ == "__main__":
    d = inputs()
    o = outputs(d)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_g2o_types.npz")
    np.savez_compressed(path, **{"in_" + k: v for k, v in d.items()}, **{"out_" + k: v for k, v in o.items()})
    print("wrote", path, os.path.getsize(path), "bytes")
