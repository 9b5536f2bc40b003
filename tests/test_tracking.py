"""Frame-to-frame pose tracking in Plücker mode (src2/stereoFrameHandler.cpp:563-853).

CPU: the oracle restatement (oracle/track_oracle.cpp) is pinned by convergence to the generating motion on noise-free data
(points only, lines only, both: a wrong Jacobian or retraction does not get there) and by its failure semantics; the kernel
logic runs from the host-emulation build against it.  GPU: the CUDA kernel through plba_track_solve against the oracle,
1e-9 on DT and the cost, on single frames, ragged batches and the degenerate cases.  The reference has no test for this
function either: parity unpinned in the task's sense."""
import numpy as np
import pytest

from pl_slam_plucker_b200 import solver, tracking as trk

CAM = (435.2, 435.2, 367.2, 252.2)


def _close(o, r):
    assert o["good"] == r["good"] and o["iters"] == r["iters"]
    np.testing.assert_allclose(r["DT"], o["DT"], rtol=0, atol=1e-9)
    if o["good"]:
        np.testing.assert_allclose(r["err"], o["err"], rtol=1e-9)
        np.testing.assert_allclose(r["DT_cov"], o["DT_cov"], rtol=1e-6, atol=1e-9 * np.abs(o["DT_cov"]).max())
    else:
        assert r["err"] == -1.0 and (r["DT_cov"] == np.eye(6)).all()


def _cases():
    out = []
    for seed in range(3):
        out.append(trk.make_frame(seed)[0])
    out.append(trk.make_frame(10, n_pt=300, n_ls=0)[0])                      # points only
    out.append(trk.make_frame(11, n_pt=0, n_ls=150)[0])                      # lines only
    F = trk.make_frame(12, n_pt=120, n_ls=60)[0]                             # inlier masks, sigma2, axis-aligned segments (overlap branches)
    rng = np.random.default_rng(5)
    F.pt_inlier = (rng.random(120) > 0.2).astype(np.uint8); F.ls_inlier = (rng.random(60) > 0.2).astype(np.uint8)
    F.ls_sigma2 = rng.uniform(0.5, 2.0, 60)
    F.ls_seg[:10, 2] = F.ls_seg[:10, 0] + 0.3                                # vertical in the previous image
    F.ls_seg[10:20, 3] = F.ls_seg[10:20, 1] + 0.3                            # horizontal
    out.append(F)
    out.append(trk.make_frame(13, n_pt=2, n_ls=0)[0])                        # too few matches: |det H| < 1 => solution_is_good = false
    out.append(trk.make_frame(14, n_pt=1000, n_ls=1000, guess_noise=0.02)[0])  # near the per-frame capacity, perturbed initial guess
    return out


def test_oracle_converges_to_the_generating_motion(oracle):
    opt = trk.Options(CAM, max_iters=10)
    for kw in (dict(n_pt=200, n_ls=0), dict(n_pt=0, n_ls=120), dict(n_pt=150, n_ls=80)):
        F, DT = trk.make_frame(21, noise=0.0, outliers=0.0, **kw)
        o = oracle.track_solve(F, opt)
        assert o["good"] == 1
        assert np.abs(o["DT"] - DT[:3]).max() < 1e-4, kw             # identity guess (5 cm, 0.02 rad off) -> the true motion; the MAD-scaled Cauchy GN converges linearly
        assert np.all(np.linalg.eigvalsh(o["DT_cov"]) > 0)           # H^-1 of a full-rank problem
    F, _ = trk.make_frame(22, n_pt=2, n_ls=0)
    o = oracle.track_solve(F, opt)
    assert o["good"] == 0 and o["err"] == -1.0 and np.abs(o["DT"] - np.eye(4)[:3]).max() == 0.0 and (o["DT_cov"] == np.eye(6)).all()   # :846-851


def _run_parity(s, oracle):
    frames = _cases()
    for iters in (5, 10):                                            # Config::maxIters / maxItersRef
        opt = trk.Options(CAM, max_iters=iters)
        res = trk.solve(s, frames, opt)                              # one batch, ragged sizes
        for F, r in zip(frames, res):
            _close(oracle.track_solve(F, opt), r)
    one = trk.solve(s, frames[:1], trk.Options(CAM))[0]              # a batch of one == the same frame inside a batch
    np.testing.assert_allclose(one["DT"], trk.solve(s, frames, trk.Options(CAM))[0]["DT"], rtol=0, atol=1e-12)   # (the block reduction's order is free)
    from pl_slam_plucker_b200.solver import LBAError
    big = trk.make_frame(1, n_pt=1030, n_ls=10)[0]
    with pytest.raises(LBAError):
        trk.solve(s, [big], trk.Options(CAM))                        # over the per-frame capacity: refused, not truncated


def test_kernel_logic_in_emulation(emu, oracle):
    s = solver.LBASolver(0, lib=emu)
    try:
        _run_parity(s, oracle)
    finally:
        s.close()


@pytest.mark.gpu
def test_gpu_tracking_parity(gpu_solver, oracle):
    _run_parity(gpu_solver, oracle)
    frames = [trk.make_frame(100 + i)[0] for i in range(300)]       # more frames than CTAs in flight
    res = trk.solve(gpu_solver, frames, trk.Options(CAM))
    for i in (0, 151, 299):
        _close(oracle.track_solve(frames[i], trk.Options(CAM)), res[i])
    assert all(r["good"] for r in res)


# ---- creation of Plücker line landmarks --------------------------------------------------------
def test_line_creation_oracle_recovers_the_3d_line(oracle):
    """Noise-free stereo segments: the triangulated NDw is the world Plücker line (unit direction) up to the sign of (n, d),
    both re-projection errors vanish and every candidate passes the sqrt(5.991) gate."""
    args, truth = trk.make_line_candidates(3, n=300, noise=0.0, outliers=0.0)
    o = oracle.create_lines(*args)
    sign = np.sign(np.sum(o["NDw"][:, 3:] * truth[:, 3:], axis=1))[:, None]
    np.testing.assert_allclose(o["NDw"] * sign, truth, atol=1e-7)
    np.testing.assert_allclose(np.linalg.norm(o["NDw"][:, 3:], axis=1), 1.0, atol=1e-12)
    assert o["err_first"].max() < 1e-7 and o["err_curr"].max() < 1e-6 and o["accept"].all()
    # NDc is a valid Plücker line: n . d = 0
    assert np.abs(np.sum(o["NDc"][:, :3] * o["NDc"][:, 3:], axis=1)).max() < 1e-9 * np.abs(o["NDc"]).max() ** 2


def _line_creation_parity(s, oracle):
    args, _ = trk.make_line_candidates(4, n=700)
    o, r = oracle.create_lines(*args), trk.create_lines(s, *args)
    for k in ("NDc", "NDw", "err_first", "err_curr"):
        np.testing.assert_allclose(r[k], o[k], rtol=1e-9, atol=1e-9)
    near = np.abs(o["err_curr"] - np.sqrt(5.991)) < 1e-9
    assert ((r["accept"] == o["accept"]) | near).all() and 0 < o["accept"].sum() < 700        # the gate is index work: bit-exact
    assert ((o["err_curr"] > np.sqrt(5.991)) == (o["accept"] == 0)).all()


def test_line_creation_in_emulation(emu, oracle):
    s = solver.LBASolver(0, lib=emu)
    try:
        _line_creation_parity(s, oracle)
    finally:
        s.close()


@pytest.mark.gpu
def test_gpu_line_creation_parity(gpu_solver, oracle):
    _line_creation_parity(gpu_solver, oracle)
