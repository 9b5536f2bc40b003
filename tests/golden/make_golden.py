"""Generates the golden fixtures in this directory from the CPU oracle (run here, committed as .npz).

The reference (kongan/PL-SLAM-plucker) cannot be built or imported in this image and ships no golden vectors, so these
pin the ORACLE's own outputs on fixed small scenes; the GPU parity tests then compare the CUDA path against them on the
GPU box, where neither /root/reference nor this script's inputs are needed.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "emu")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_lib                      # only used for its scene generator entry points
from helpers import robust_prefix
from oracle import loader as orc
from pl_slam_plucker_b200 import abi, scene

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"g_faithful": (abi.PROFILE_G, 0, 0, 0), "g_fixed": (abi.PROFILE_G, 1, 0, 0), "h_end_faithful": (abi.PROFILE_H_END, 0, 1, 0),
         "h_plk_fixed": (abi.PROFILE_H_PLK, 1, 0, 0),
         "gba_faithful": (abi.PROFILE_H_END, 0, 1, abi.SHELL_GBA), "gba_fixed": (abi.PROFILE_H_END, 1, 1, abi.SHELL_GBA)}   # Global BA shell

if __name__ == "__main__":
    L = emu_lib.load()
    only = sys.argv[1:]
    for name, (prof, q, line_mode, shell) in CASES.items():
        if only and name not in only:
            continue
        if shell == abi.SHELL_GBA:          # every KF but KF 0 free, endpoint lines included
            P = scene.make_scene(1, lib=L, n_kf_free=6, n_kf_fixed=1, n_pt=120, n_ls=30, line_mode=line_mode, seed=4243)
        else:
            P = scene.make_scene(1, lib=L, n_kf_free=5, n_kf_fixed=2, n_pt=120, n_ls=0 if prof == abi.PROFILE_H_END else 40, line_mode=line_mode, seed=4242)
        r = orc.solve(P, abi.Options(prof, q, shell=shell))
        n = robust_prefix(r.trace) if prof == abi.PROFILE_G else len(r.trace)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), profile=prof, quirks=q, shell=shell, cam=P.cam, kf_T_wc=P.kf_T_wc, kf_slot=P.kf_slot,
                            pt_xyz=P.pt_xyz, po_lm=P.po_lm, po_kf=P.po_kf, po_uv=P.po_uv, ls_plk=P.ls_plk, ls_end=P.ls_end, lo_lm=P.lo_lm,
                            lo_kf=P.lo_kf, lo_ab=P.lo_ab, x_pose=P.x_pose, n_robust=n, trace_chi=r.trace["chi"], trace_chi_new=r.trace["chi_new"],
                            trace_lambda=r.trace["lambda"], trace_accepted=r.trace["accepted"], out_kf_T_wc=r.kf_T_wc, out_pt_xyz=r.pt_xyz,
                            out_ls_orth=r.ls_orth, out_ls_plk=r.ls_plk, out_ls_end=r.ls_end, out_po_flags=r.po_flags, out_lo_flags=r.lo_flags)
        print(name, "trace", len(r.trace), "robust prefix", n)
