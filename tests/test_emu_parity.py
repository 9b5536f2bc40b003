"""CPU-side rehearsal of the GPU parity tests: the SAME kernel sources compiled with -DPLBA_HOST_EMU (tests/emu, test
tooling only) run the small parity cases against the oracle, so that kernel logic is checked on a box without a GPU.
The product library (libplba.so) is never replaced by this build; `-m gpu` runs the real thing."""
import pytest

import test_gpu_parity as g
from pl_slam_plucker_b200 import solver


@pytest.fixture(scope="module", params=[1, 2], ids=["chunk-kernels", "warp-kernels"])
def gpu_solver(emu, request):
    """Both implementations of the assembly / update stage (CTA-chunk and warp-autonomous) run every case."""
    s = solver.LBASolver(0, lib=emu)
    s.set_kernel_path(request.param)
    yield s
    s.close()


test_small_window_all_profiles = g.test_small_window_all_profiles
test_golden_fixtures = g.test_golden_fixtures
test_large_window_solvers = g.test_large_window_solvers
test_largest_single_cta_window = g.test_largest_single_cta_window
test_loop_closure_shaped_window = g.test_loop_closure_shaped_window
test_batch_equals_individual_solves = g.test_batch_equals_individual_solves
test_reduced_system_blocks_and_sparsity = g.test_reduced_system_blocks_and_sparsity
test_edge_cases = g.test_edge_cases
test_sigma_weights = g.test_sigma_weights
test_global_ba_shell = g.test_global_ba_shell
test_both_kernel_paths = g.test_both_kernel_paths
test_randomised_window_shapes = g.test_randomised_window_shapes
test_map_handler_interface = g.test_map_handler_interface
test_randomised_large_windows = g.test_randomised_large_windows
test_all_keyframes_fixed_hand_lm = g.test_all_keyframes_fixed_hand_lm
test_randomised_large_windows_g_faithful = g.test_randomised_large_windows_g_faithful
test_numeric_failure_is_reported = g.test_numeric_failure_is_reported


@pytest.mark.parametrize("env", [dict(PLBA_DENSE_GROUP="1"), dict(PLBA_DENSE_GROUP="3"), dict(PLBA_DENSE_GROUP="4"), dict(PLBA_DENSE_K1="1")],
                         ids=["group1", "group3", "group4", "k1"])
def test_dense_cholesky_launch_geometry(emu, oracle, env):
    """The dense tiled Cholesky's launch sequence (panel groups, in-group row updates, the split trailing update) walked on the CPU for
    every group size: a 68-keyframe window with loop closures (order 408: four full panels and one of 24), profile G fixed, against the oracle."""
    import os
    from pl_slam_plucker_b200 import abi, scene
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        s = solver.LBASolver(0, lib=emu)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        P = scene.make_scene(1, n_kf_free=68, n_kf_fixed=2, n_pt=1800, n_ls=300, loop_every=30, seed=61)
        opt = abi.Options(abi.PROFILE_G, 1, iters_stage1=2, iters_stage2=1)
        r, o = s.solve(P, opt), oracle.solve(P, opt)
        assert s.kernel_path()["solver"] == "dense-dmma"
        g.assert_trace_close(o.trace, r.trace, abi.PROFILE_G)
        g.assert_state_close(o, r, P, abi.PROFILE_G)
    finally:
        s.close()


def _with_threads(n, fn):
    import os
    old = os.environ.get("PLBA_HOST_THREADS")
    os.environ["PLBA_HOST_THREADS"] = str(n)
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("PLBA_HOST_THREADS", None)
        else:
            os.environ["PLBA_HOST_THREADS"] = old


def _same_results(a, b):
    import numpy as np
    for f in ("kf_T_wc", "pt_xyz", "ls_plk", "ls_orth", "po_chi2", "lo_chi2", "po_flags", "lo_flags"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.n_trials == b.n_trials and np.array_equal(a.trace, b.trace)


@pytest.mark.parametrize("path", [1, 2], ids=["chunk-kernels", "warp-kernels"])
def test_host_preparation_parallel_equals_serial(emu, path):
    """The host side of plba_upload / plba_download on several threads (a large window: slice-parallel signature order, counting sort, copies; a
    batch: window-parallel validation, runs, layout statistics) lays the problem out exactly as one thread does: same layout statistics, and —
    the emulation adds in a fixed order — bit-identical results, chi2 and flags in the caller's order included."""
    from pl_slam_plucker_b200 import abi, scene
    opt = abi.Options(abi.PROFILE_G, 1, iters_stage1=1, iters_stage2=0)
    big = scene.make_scene(1, n_kf_free=30, n_kf_fixed=2, n_pt=60000, n_ls=56000, seed=77)      # > 200 000 observations and > 40 000 landmarks in BOTH classes: the parallel loops are taken (lines: the hash-grouping path)
    assert big.n_pobs > 200000 and big.n_lobs > 200000
    batch = scene.make_batch(12, 3, n_pt=300, n_ls=60)
    out = {}
    for nt in (1, 4):
        s = solver.LBASolver(0, lib=emu)
        s.set_kernel_path(path)
        try:
            def run():
                r = s.solve(big, opt)
                st = s.layout_stats()
                rc, rs = s.solve_batch(batch, opt)
                return r, st, rs, s.layout_stats()
            out[nt] = _with_threads(nt, run)
        finally:
            s.close()
    (r1, st1, b1, bst1), (r4, st4, b4, bst4) = out[1], out[4]
    assert st1 == st4 and bst1 == bst4
    _same_results(r1, r4)
    for x, y in zip(b1, b4):
        _same_results(x, y)
