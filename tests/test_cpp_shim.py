"""The C++ host shim (csrc/shim/map_handler_shim.h: the reference's own entry-point names over the C ABI) compiled with
plain g++ and run as the reference's host code would run it.  CPU: linked against the kernel-emulation build (logic check);
GPU: linked against the product library."""
import os
import re
import subprocess

import numpy as np
import pytest

from pl_slam_plucker_b200 import abi, scene

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_demo.cpp")


def _build(tmp_path, libdir, libname):
    exe = str(tmp_path / "shim_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", exe, SRC, "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir])
    return exe


def _check(out, oracle):
    P = scene.make_scene(1, n_kf_free=5, n_kf_fixed=2, n_pt=120, n_ls=30, seed=13)
    o = oracle.solve(P, abi.Options(abi.PROFILE_G, 0))
    m = re.search(r"g2o_path trials=(\d+) chi_first=(\S+) chi_last=(\S+) bad_pt_obs=(\d+) bad_ls_obs=(\d+) T2=(\S+),(\S+),(\S+)", out)
    assert m, out
    assert abs(float(m.group(2)) - o.trace["chi"][0]) <= 1e-9 * o.trace["chi"][0]
    assert int(m.group(4)) == int(((o.po_flags & abi.OBS_BAD) != 0).sum()) and int(m.group(5)) == int(((o.lo_flags & abi.OBS_BAD) != 0).sum())
    np.testing.assert_allclose([float(m.group(i)) for i in (6, 7, 8)], o.kf_T_wc[2][[3, 7, 11]], atol=1e-8)
    assert re.search(r"hand_lm rc=0 iters=\d+", out) and "discarded rc=-1" in out
    m2 = re.search(r"culled=(\d+) expected=(\d+) points_left=(\d+)", out)      # removeBadMapLandmarksForPluker after the LBA
    assert m2 and m2.group(1) == m2.group(2) and int(m2.group(1)) > 0 and int(m2.group(3)) < 120
    assert "gba iters=15" in out
    # StereoFrameHandlerShim: the pose tracker recovers the pure translation (0.02, 0, 0.05); the triangulated line has the segment's direction
    mt = re.search(r"track good=1 iters=\d+ t=(\S+),(\S+),(\S+) err=(\S+)", out)
    assert mt, out
    np.testing.assert_allclose([float(mt.group(i)) for i in (1, 2, 3)], [0.02, 0.0, 0.05], atol=1e-5)
    ml = re.search(r"newline accepted=1 error=(\S+) error2=(\S+) d=(\S+),(\S+),(\S+)", out)
    assert ml and float(ml.group(1)) < 1e-6 and float(ml.group(2)) < 1e-6, out
    dtrue = np.array([1.1, -0.5, 1.0]) / np.linalg.norm([1.1, -0.5, 1.0]); dgot = np.array([float(ml.group(i)) for i in (3, 4, 5)])
    assert min(np.abs(dgot - dtrue).max(), np.abs(dgot + dtrue).max()) < 1e-6             # faithful GBA: err /= 0 in every pass, so all maxItersLba passes run (src/mapHandler.cpp:3662-3665)


def test_shim_compiles_and_runs_on_the_emulation_build(tmp_path, oracle, emu):
    import emu_lib
    exe = _build(tmp_path, os.path.dirname(emu_lib.SO), os.path.basename(emu_lib.SO))
    _check(subprocess.check_output([exe], text=True), oracle)


def test_shim_fails_loudly_without_cuda(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    exe = _build(tmp_path, os.path.join(ROOT, "pl_slam_plucker_b200"), "libplba.so")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_shim_on_the_gpu(tmp_path, oracle):
    exe = _build(tmp_path, os.path.join(ROOT, "pl_slam_plucker_b200"), "libplba.so")
    _check(subprocess.check_output([exe], text=True), oracle)
