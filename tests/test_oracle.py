"""CPU tests of the oracle itself.  The reference has no tests or golden vectors for its LBA, so the
oracle is pinned by: finite-difference checks of every Jacobian against the reference's own retractions, round trips,
Schur == literal dense full-system solve, and committed golden fixtures of its own outputs (regression pin)."""
import os

import numpy as np
import pytest

from pl_slam_plucker_b200 import abi, scene

CAM = np.array([435.2, 435.2, 367.2, 252.2])
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rand_T(oracle, rng):
    return oracle.expmap_se3(np.r_[rng.normal(size=3) * 0.5, rng.normal(size=3) * 0.3])


def test_se3_exp_log_roundtrip(oracle):
    rng = np.random.default_rng(1)
    for _ in range(20):
        x = np.r_[rng.normal(size=3), rng.normal(size=3) * 0.4]
        T = oracle.expmap_se3(x)
        np.testing.assert_allclose(oracle.logmap_se3(T), x, atol=1e-11)
        np.testing.assert_allclose(oracle.inverse_se3(T) @ T, np.eye(4), atol=1e-13)
    # small-angle branch (theta < 1e-6): R = I, t = rho (src2/auxiliar.cpp:131-132)
    T = oracle.expmap_se3([1, 2, 3, 1e-8, 0, 0])
    np.testing.assert_allclose(T[:3, :3], np.eye(3)); np.testing.assert_allclose(T[:3, 3], [1, 2, 3])


def test_orth_pluker_roundtrip(oracle):
    rng = np.random.default_rng(2)
    for _ in range(20):
        P, Q = rng.normal(size=3) * 3 + [0, 0, 6], rng.normal(size=3) * 3 + [0, 0, 6]
        d = (Q - P) / np.linalg.norm(Q - P); n = np.cross(P, d)
        pl = np.r_[n, d]
        o = oracle.pluker_to_orth(pl)
        back = oracle.orth_to_pluker(o)
        np.testing.assert_allclose(back * np.linalg.norm(pl), pl, atol=1e-12)     # changeOrthToPluker returns the unit-norm 6-vector
        np.testing.assert_allclose(oracle.update_orth(o, np.zeros(4)), o, atol=1e-12)


def test_point_edge_jacobians_fd(oracle):
    rng = np.random.default_rng(3)
    h = 1e-6
    for _ in range(5):
        T = _rand_T(oracle, rng); Pw = np.array([0.3, -0.2, 5.0]) + rng.normal(size=3) * 0.3; obs = np.array([400.0, 260.0])
        e, Ji, Jj = oracle.point_edge(CAM, T, Pw, obs)
        for k in range(6):
            d = np.zeros(6); d[k] = h
            fd = (oracle.point_edge(CAM, oracle.pose_oplus(T, d), Pw, obs)[0] - oracle.point_edge(CAM, oracle.pose_oplus(T, -d), Pw, obs)[0]) / (2 * h)
            np.testing.assert_allclose(Jj[:, k], fd, rtol=1e-6, atol=1e-6)
        for k in range(3):
            d = np.zeros(3); d[k] = h
            fd = (oracle.point_edge(CAM, T, Pw + d, obs)[0] - oracle.point_edge(CAM, T, Pw - d, obs)[0]) / (2 * h)
            np.testing.assert_allclose(Ji[:, k], fd, rtol=1e-6, atol=1e-6)


def test_line_edge_jacobians_fd(oracle):
    """Landmark Jacobian is FD-exact; the pose Jacobian is FD-exact only with Q12 fixed."""
    rng = np.random.default_rng(4)
    h = 1e-6
    for _ in range(5):
        T = _rand_T(oracle, rng)
        P, Q = np.array([0.5, 0.2, 6.0]) + rng.normal(size=3) * 0.3, np.array([-0.4, 0.5, 7.0]) + rng.normal(size=3) * 0.3
        d = (Q - P) / np.linalg.norm(Q - P); orth = oracle.pluker_to_orth(np.r_[np.cross(P, d), d])
        obs = np.array([300.0, 200.0, 420.0, 280.0])
        e, Ji, Jj = oracle.line_edge(CAM, T, orth, obs, faithful=False)
        _, Ji_f, Jj_f = oracle.line_edge(CAM, T, orth, obs, faithful=True)
        np.testing.assert_allclose(Ji_f, Ji)
        fdj = np.zeros((2, 6)); fdi = np.zeros((2, 4))
        for k in range(6):
            dd = np.zeros(6); dd[k] = h
            fdj[:, k] = (oracle.line_edge(CAM, oracle.pose_oplus(T, dd), orth, obs)[0] - oracle.line_edge(CAM, oracle.pose_oplus(T, -dd), orth, obs)[0]) / (2 * h)
        for k in range(4):
            dd = np.zeros(4); dd[k] = h
            fdi[:, k] = (oracle.line_edge(CAM, T, oracle.update_orth(orth, dd), obs)[0] - oracle.line_edge(CAM, T, oracle.update_orth(orth, -dd), obs)[0]) / (2 * h)
        np.testing.assert_allclose(Jj, fdj, rtol=2e-6, atol=1e-5)
        np.testing.assert_allclose(Ji, fdi, rtol=2e-6, atol=1e-5)
        assert np.abs(Jj_f - fdj).max() > 1e-2 * np.abs(fdj).max()      # the faithful pose Jacobian is NOT the derivative


def test_h_point_term_is_minus_gradient(oracle):
    """Profile H: J_p = -d|e|/d xi for xi left-perturbing T_cw (translation first); J_l = -d|e|/dX (a12)."""
    rng = np.random.default_rng(5)
    h = 1e-6
    T = _rand_T(oracle, rng); X = np.array([0.4, -0.3, 6.0]); obs = np.array([390.0, 240.0])
    Jp, Jl, r, w = oracle.h_term(0, CAM, T, X, obs)
    assert abs(w - 1.0 / (1.0 + r * r)) < 1e-15
    for k in range(3):
        d = np.zeros(3); d[k] = h
        fd = (oracle.h_term(0, CAM, T, X + d, obs)[2] - oracle.h_term(0, CAM, T, X - d, obs)[2]) / (2 * h)
        assert abs(Jl[k] + fd) < 1e-5 * max(1.0, abs(fd))
    for k in range(6):
        d = np.zeros(6); d[k] = h
        Tp, Tm = oracle.expmap_se3(d) @ T, oracle.expmap_se3(-d) @ T
        fd = (oracle.h_term(0, CAM, Tp, X, obs)[2] - oracle.h_term(0, CAM, Tm, X, obs)[2]) / (2 * h)
        assert abs(Jp[k] + fd) < 1e-5 * max(1.0, abs(fd))


@pytest.mark.parametrize("profile,quirks", [(abi.PROFILE_G, 0), (abi.PROFILE_G, 1), (abi.PROFILE_H_END, 0), (abi.PROFILE_H_END, 1), (abi.PROFILE_H_PLK, 1)])
def test_schur_equals_dense_full_system(oracle, emu, profile, quirks):
    """Landmark Schur + reduced solve == the literal dense N x N solve of the hand LM (src/mapHandler.cpp:2343,2566-2568)."""
    P = scene.make_scene(1, lib=emu, n_kf_free=4, n_kf_fixed=2, n_pt=60, n_ls=20, line_mode=1 if profile == abi.PROFILE_H_END else 0)
    opt = abi.Options(profile, quirks)
    a, b = oracle.solve(P, opt), oracle.solve(P, opt, dense=True)
    assert len(a.trace) == len(b.trace)
    fin = np.isfinite(a.trace["chi"])
    np.testing.assert_allclose(a.trace["chi"][fin], b.trace["chi"][fin], rtol=1e-9)
    np.testing.assert_allclose(a.kf_T_wc, b.kf_T_wc, atol=1e-10)
    np.testing.assert_allclose(a.pt_xyz, b.pt_xyz, atol=1e-9)


def test_profile_g_lm_gain_ratio_is_one_for_exact_jacobians(oracle, emu):
    """Points only: exact Jacobians => the g2o gain ratio of the gated (non-robust) stage stays ~1 (model == cost)."""
    P = scene.make_scene(1, lib=emu, n_ls=0, n_pt=400, outlier_frac=0.0)
    r = oracle.solve(P, abi.Options(abi.PROFILE_G, 1))
    st2 = r.trace[r.trace["stage"] == 1]
    assert len(st2) == 10 and (st2["accepted"] == 1).all()
    assert np.all(np.abs(st2["rho"][:6] - 1.0) < 0.05)      # later trials sit at the noise floor (chi change ~ rounding)
    assert np.all(np.abs(st2["rho"] - 1.0) < 0.25)


def test_profile_g_converges_to_truth(oracle, emu):
    P, truth = scene.make_scene(1, lib=emu, with_truth=True, n_pt=600, n_ls=150, pixel_noise=0.0, outlier_frac=0.0)
    r = oracle.solve(P, abi.Options(abi.PROFILE_G, 1))
    assert np.abs(r.kf_T_wc - truth["kf_T_wc"]).max() < np.abs(P.kf_T_wc - truth["kf_T_wc"]).max() * 0.05
    assert r.trace["chi"][-1] < 1e-3 * r.trace["chi"][0]


def test_empty_problem_is_discarded(oracle, emu):
    P = scene.make_scene(1, lib=emu, n_kf_free=3, n_kf_fixed=1, n_pt=5, n_ls=0)
    E = abi.Problem(P.cam, P.kf_T_wc, P.kf_slot, np.zeros((0, 3)), None, None, None)
    assert oracle.solve(E, abi.Options(abi.PROFILE_H_END)).rc == abi.DISCARDED      # src/mapHandler.cpp:1499-1500


def test_h_faithful_first_error_is_infinite(oracle, emu):
    P = scene.make_scene(1, lib=emu, n_kf_free=4, n_kf_fixed=2, n_pt=80, n_ls=0)
    r = oracle.solve(P, abi.Options(abi.PROFILE_H_END, 0))
    assert np.isinf(r.trace["chi"][0]) and np.isfinite(r.trace["chi"][1])           # Q1: err /= 0 after pass 0 only
    assert r.trace["accepted"][1] == 1                                              # inf as err_prev => first loop step applied


@pytest.mark.parametrize("quirks", [0, 1])
def test_global_ba_shell(oracle, emu, quirks):
    """levMarquardtOptimizationGBA (src/mapHandler.cpp:3128-3728): Schur == literal dense solve inside the GBA shell too;
    faithful: lambda0 = 1e-5 * trunc(Hmax) (`int Hmax`, :3386-3392), cost x/0 in every pass => all passes run, all applied."""
    P = scene.make_scene(1, lib=emu, n_kf_free=6, n_kf_fixed=1, n_pt=80, n_ls=20, line_mode=1, seed=77)
    opt = abi.Options(abi.PROFILE_H_END, quirks, shell=abi.SHELL_GBA)
    a, b = oracle.solve(P, opt), oracle.solve(P, opt, dense=True)
    assert len(a.trace) == len(b.trace)
    np.testing.assert_allclose(a.kf_T_wc, b.kf_T_wc, atol=1e-10)
    np.testing.assert_allclose(a.pt_xyz, b.pt_xyz, atol=1e-9)
    assert a.pt_inlier.all() and a.ls_inlier.all()                     # no `inlier` rule (:3705-3726)
    lba = oracle.solve(P, abi.Options(abi.PROFILE_H_END, quirks))
    if quirks == 0:
        assert len(a.trace) == opt.max_iters_lba and np.isinf(a.trace["chi"]).all() and (a.trace["accepted"][1:] == 1).all()
        lam_g, lam_l = a.trace["lambda"][1], lba.trace["lambda"][1]      # lambda in force for the first loop pass
        hmax = lam_l / 1e-5
        np.testing.assert_allclose(lam_g, 1e-5 * np.floor(hmax), rtol=1e-12)
    else:
        # intended maths: finite cost normalised by the observation count; epsilon() thresholds => runs longer than the LBA shell
        assert np.isfinite(a.trace["chi"]).all() and len(a.trace) >= len(lba.trace)
        np.testing.assert_allclose(a.trace["chi"][0] * P.n_obs, lba.trace["chi"][0] * (P.n_pt + P.n_ls), rtol=1e-12)


@pytest.mark.parametrize("name", ["g_faithful", "g_fixed", "h_end_faithful", "h_plk_fixed", "gba_faithful", "gba_fixed"])
def test_golden_fixtures(oracle, name):
    """Committed oracle outputs (tests/golden/make_golden.py): guards the oracle against silent drift."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    prob = abi.Problem(z["cam"], z["kf_T_wc"], z["kf_slot"], z["pt_xyz"], z["po_lm"], z["po_kf"], z["po_uv"], ls_plk=z["ls_plk"],
                       ls_end=z["ls_end"], lo_lm=z["lo_lm"], lo_kf=z["lo_kf"], lo_ab=z["lo_ab"], x_pose=z["x_pose"])
    r = oracle.solve(prob, abi.Options(int(z["profile"]), int(z["quirks"]), shell=int(z["shell"]) if "shell" in z else abi.SHELL_LBA))
    n = int(z["n_robust"])
    fin = np.isfinite(z["trace_chi"][:n])
    np.testing.assert_allclose(r.trace["chi"][:n][fin], z["trace_chi"][:n][fin], rtol=1e-9)
    assert (r.trace["accepted"][:n] == z["trace_accepted"][:n]).all()
    np.testing.assert_allclose(r.kf_T_wc, z["out_kf_T_wc"], atol=1e-8)
    np.testing.assert_allclose(r.pt_xyz, z["out_pt_xyz"], atol=1e-8)


def test_culling_after_lba_host_bookkeeping():
    """removeBadMapLandmarksForPluker (src/mapHandler.cpp:3816-3897) in the Python mirror: pure index work, no solver."""
    from pl_slam_plucker_b200 import map_handler as mhm
    mh = mhm.MapHandler(mhm.PinholeStereoCamera(400, 400, 320, 240), solver=None)
    for k in range(14):
        mh.map_keyframes.append(mhm.KeyFrame(k, np.eye(4)))
    mh.max_kf_idx = 13
    def point(idx, base, n_obs, inlier=True, local=False):
        p = mhm.MapPoint(idx, [0, 0, 5], local=local); p.inlier = inlier
        for j in range(n_obs):
            p.addMapPointObservation(base + j, [10.0, 20.0])
        mh.map_points.append(p); mh.map_points_kf_idx.setdefault(base, []).append(idx); mh.map_keyframes[base].stereo_pt_idx.append(idx)
        return p
    point(0, 0, 6)                         # old, enough observations, inlier: kept
    point(1, 0, 3)                         # old, too few observations (< minLMObs = 5): culled
    point(2, 1, 6, inlier=False)           # old outlier (the LBA write-back cleared `inlier`): culled
    point(3, 5, 2)                         # first seen 8 keyframes ago: too recent to be judged
    point(4, 0, 2, local=True)             # still local: kept
    ln = mhm.MapLine(0, NDw=[0, 0, 1, 1, 0, 0], local=False)
    for j in range(2):
        ln.addMapLineObservation(j, [1.0, 2.0, 3.0, 4.0])
    mh.map_lines.append(ln); mh.map_lines_kf_idx[0] = [0]; mh.map_keyframes[0].stereo_ls_idx.append(0)
    assert mh.removeBadMapLandmarksForPluker() == (2, 1)
    assert [p is None for p in mh.map_points] == [False, True, True, False, False] and mh.map_lines[0] is None
    assert mh.map_points_kf_idx[0] == [0, 4] and mh.map_points_kf_idx[1] == [] and mh.map_lines_kf_idx[0] == []
    assert mh.map_keyframes[0].stereo_pt_idx == [0, -1, 4] and mh.map_keyframes[1].stereo_pt_idx == [-1] and mh.map_keyframes[0].stereo_ls_idx == [-1]
    assert mh.removeBadMapLandmarksForPluker() == (0, 0)          # idempotent
