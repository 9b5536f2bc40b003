"""Multi-GPU host logic on CPU: world_size-2 `gloo` run of the landmark-sharded LBA.

Each rank holds one landmark shard, the reduced camera system is all-reduced through the library's exchange hook, and the
result must equal the single-process oracle on the whole window.  The kernels run from the host-emulation build (tests/emu,
test tooling); on the GPU box the same driver runs over NCCL (bench.py --gpus N, tests/test_gpu_parity.py covers the shard
linearity on one GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, quirks, out_dir, shape, empty_rank=-1):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import emu_lib
    from pl_slam_plucker_b200 import abi, scene, sharded, solver
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        P = scene.make_scene(1, seed=31, **shape)
        s = solver.LBASolver(0, lib=emu_lib.load())
        sh = sharded.ShardedLBA(s, rank, world)
        opt = abi.Options(abi.PROFILE_G, quirks)
        masks = None
        if empty_rank >= 0:      # one rank holds every landmark, the other an EMPTY shard (it still takes part in every exchange)
            own = np.ones if rank != empty_rank else np.zeros
            masks = (own(P.n_pt, bool), own(P.n_ls, bool))
        r = sh.solve(P, opt, masks)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), kf_T_wc=r.kf_T_wc, pt_xyz=r.pt_xyz, ls_orth=r.ls_orth, ls_plk=r.ls_plk,
                 pt_index=sh.pt_index, ls_index=sh.ls_index, chi=r.trace["chi"], chi_new=r.trace["chi_new"], lam=r.trace["lambda"],
                 accepted=r.trace["accepted"], rho=r.trace["rho"], po_flags=r.po_flags, lo_flags=r.lo_flags,
                 po_sel=(masks or sharded.shard_masks(P, rank, world))[0][P.po_lm], lo_sel=(masks or sharded.shard_masks(P, rank, world))[1][P.lo_lm],
                 status=r.status, n_trials=r.n_trials)
        s.close()
    finally:
        dist.destroy_process_group()


SMALL = dict(n_kf_free=8, n_kf_fixed=2, n_pt=300, n_ls=80)          # reduced camera system solved in one CTA: the whole [S | g] is exchanged
LARGE = dict(n_kf_free=40, n_kf_fixed=2, n_pt=800, n_ls=200)        # block cyclic reduction: only the band [D | U | b] of the nodes is exchanged,
                                                                    # and the ranks first agree on the band width of the summed system


@pytest.mark.parametrize("quirks,shape", [(0, SMALL), (1, SMALL), (1, LARGE)], ids=["faithful-small", "fixed-small", "fixed-large-banded"])
def test_two_rank_sharded_lba_equals_single_process_oracle(tmp_path, oracle, quirks, shape):
    from helpers import COST_RTOL, RHO_MARGIN, STATE_ATOL
    from pl_slam_plucker_b200 import abi, scene
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), quirks, str(tmp_path), shape), nprocs=world, join=True)
    P = scene.make_scene(1, seed=31, **shape)
    o = oracle.solve(P, abi.Options(abi.PROFILE_G, quirks))
    z = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # every rank took the same LM decisions and ended with the same poses
    for k in ("chi", "chi_new", "lam", "accepted"):
        np.testing.assert_array_equal(z[0][k], z[1][k])
    np.testing.assert_array_equal(z[0]["kf_T_wc"], z[1]["kf_T_wc"])
    n = 0
    for t in o.trace:
        if abs(t["rho"]) < RHO_MARGIN and t["chi_new"] != 0.0:
            break
        n += 1
    n = min(n, len(z[0]["chi"]))
    assert n >= 5
    np.testing.assert_allclose(z[0]["chi"][:n], o.trace["chi"][:n], rtol=COST_RTOL)
    np.testing.assert_allclose(z[0]["chi_new"][:n], o.trace["chi_new"][:n], rtol=COST_RTOL)
    assert (z[0]["accepted"][:n] == o.trace["accepted"][:n]).all()
    np.testing.assert_allclose(z[0]["kf_T_wc"], o.kf_T_wc, atol=STATE_ATOL)
    # the shards partition the landmarks; together they reproduce the oracle's landmarks and observation flags
    pt = np.full_like(o.pt_xyz, np.nan); ls = np.full_like(o.ls_orth, np.nan)
    pf = np.full(P.n_pobs, 255, np.uint8); lf = np.full(P.n_lobs, 255, np.uint8)
    for r in range(world):
        pt[z[r]["pt_index"]] = z[r]["pt_xyz"]; ls[z[r]["ls_index"]] = z[r]["ls_orth"]
        pf[z[r]["po_sel"]] = z[r]["po_flags"]; lf[z[r]["lo_sel"]] = z[r]["lo_flags"]
    assert sorted(np.r_[z[0]["pt_index"], z[1]["pt_index"]].tolist()) == list(range(P.n_pt))
    np.testing.assert_allclose(pt, o.pt_xyz, atol=STATE_ATOL)
    np.testing.assert_allclose(ls, o.ls_orth, atol=STATE_ATOL)
    near = np.abs(o.po_chi2 - 5.991) < 1e-6
    assert ((pf == o.po_flags) | near).all()
    nearl = np.abs(o.lo_chi2 - 5.991) < 1e-6
    assert ((lf == o.lo_flags) | nearl).all()


@pytest.mark.parametrize("shape", [SMALL, LARGE], ids=["small", "large-banded"])
def test_rank_with_empty_shard_takes_part(tmp_path, oracle, shape):
    """A rank whose shard holds no landmark at all (more ranks than a window can feed) must not drop out: it contributes zeros to every
    exchange, solves the same summed system and ends with the same poses as the rank that holds everything — which equal the oracle's."""
    from helpers import COST_RTOL, RHO_MARGIN, STATE_ATOL
    from pl_slam_plucker_b200 import abi, scene
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 1, str(tmp_path), shape, 1), nprocs=world, join=True)
    P = scene.make_scene(1, seed=31, **shape)
    o = oracle.solve(P, abi.Options(abi.PROFILE_G, 1))
    z = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert int(z[0]["status"]) == 0 and int(z[1]["status"]) == 0 and int(z[0]["n_trials"]) == int(z[1]["n_trials"]) > 0
    for k in ("chi", "chi_new", "lam", "accepted"):
        np.testing.assert_array_equal(z[0][k], z[1][k])
    np.testing.assert_array_equal(z[0]["kf_T_wc"], z[1]["kf_T_wc"])
    assert z[1]["pt_index"].size == 0 and z[1]["po_flags"].size == 0
    n = 0
    for t in o.trace:
        if abs(t["rho"]) < RHO_MARGIN and t["chi_new"] != 0.0:
            break
        n += 1
    n = min(n, len(z[0]["chi"]))
    assert n >= 5
    np.testing.assert_allclose(z[0]["chi"][:n], o.trace["chi"][:n], rtol=COST_RTOL)
    np.testing.assert_allclose(z[0]["kf_T_wc"], o.kf_T_wc, atol=STATE_ATOL)
    np.testing.assert_allclose(z[0]["pt_xyz"], o.pt_xyz, atol=STATE_ATOL)


def test_shard_partition_properties():
    from pl_slam_plucker_b200 import scene, sharded
    P = scene.make_scene(4, n_pt=4000, n_ls=1000)
    for world in (2, 4, 8):
        ranges = sharded.kf_ranges(P, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == P.n_kf
        assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
        masks = [sharded.shard_masks(P, r, world) for r in range(world)]
        assert (np.sum([m[0] for m in masks], axis=0) == 1).all() and (np.sum([m[1] for m in masks], axis=0) == 1).all()
        obs = [int(m[0][P.po_lm].sum() + m[1][P.lo_lm].sum()) for m in masks]
        assert min(obs) > 0 and max(obs) < 2.0 * np.mean(obs)            # balanced by observation count
        sh, pi, li = sharded.shard_problem(P, 1, world)
        assert sh.n_kf == P.n_kf and sh.n_pt == pi.size and (np.diff(sh.po_lm) >= 0).all()
