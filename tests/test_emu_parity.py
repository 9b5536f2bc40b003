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
