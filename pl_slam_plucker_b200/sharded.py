"""Landmark-sharded LBA over the GPUs of one box (BASELINE configs 4-5).

Partition: every landmark, with ALL its observations, belongs to the rank that owns its base keyframe (the first KF that
observes it — the reference's own notion, `map_points_kf_idx`, include/mapHandler.h:148-149).  Keyframes are cut into
contiguous ranges balanced by observation count; poses are replicated.  Per LM trial the path has exactly one exchange
step: the reduced camera system [S | g | diag(H_pp)] and the cost scalars are summed over ranks (NCCL all-reduce over
NVLink on the GPUs; gloo in the CPU tests), after which every rank solves the same small system, back-substitutes its own
landmarks and takes the same LM decision.  There is no other data-path collective.

On GPUs the collective lives INSIDE the library: `ShardedLBA(..., nccl=True)` has rank 0 create an NCCL unique id
(plba_comm_unique_id), ships it to the other ranks over torch.distributed (plumbing: a 128-byte broadcast, any backend) and
every rank calls plba_comm_init_rank(); from then on the library itself issues ncclAllReduce on its stream, with no host
synchronisation inside the LM loop — exactly what a C++ host gets.  `nccl=False` hands the library a
caller-supplied all-reduce instead (plba_set_allreduce): that is how the CPU tests run the same protocol over gloo.
"""
import numpy as np

from . import abi


def base_keyframes(P):
    """Row of the first observing KF per landmark (points, lines); landmarks without observations go to KF 0."""
    home_p = np.zeros(P.n_pt, np.int64)
    if P.n_pobs:
        home_p[P.po_lm[::-1]] = P.po_kf[::-1]          # reversed assignment leaves the FIRST observation of each landmark
    home_l = np.zeros(P.n_ls, np.int64)
    if P.n_lobs:
        home_l[P.lo_lm[::-1]] = P.lo_kf[::-1]
    return home_p, home_l


def kf_ranges(P, world):
    """Contiguous KF ranges [lo, hi) per rank, balanced by the observations of the landmarks based there."""
    home_p, home_l = base_keyframes(P)
    w = np.zeros(P.n_kf, np.float64)
    if P.n_pobs:
        np.add.at(w, home_p[P.po_lm], 1.0)
    if P.n_lobs:
        np.add.at(w, home_l[P.lo_lm], 1.5)             # a line observation costs about 1.5 point observations
    cum = np.cumsum(w)
    total = cum[-1] if cum.size else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world)) + 1 if total > 0 else 0)
    cuts.append(P.n_kf)
    cuts = np.maximum.accumulate(np.minimum(cuts, P.n_kf))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def shard_masks(P, rank, world):
    lo, hi = kf_ranges(P, world)[rank]
    home_p, home_l = base_keyframes(P)
    return (home_p >= lo) & (home_p < hi), (home_l >= lo) & (home_l < hi)


def shard_problem(P, rank, world):
    """The rank's shard: all keyframes, its landmarks re-indexed densely; returns (shard, pt_index, ls_index)."""
    mp, ml = shard_masks(P, rank, world)
    return P.subset_landmarks(mp, ml), np.flatnonzero(mp), np.flatnonzero(ml)


class _DevBuf:
    """Zero-copy view of library-owned device memory for torch (`__cuda_array_interface__`)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def make_allreduce(group=None, device=None):
    """fn(ptr, n, stream) for LBASolver.set_allreduce: in-place sum (n > 0) or max (n < 0) of n doubles over `group`."""
    import ctypes
    import torch
    import torch.distributed as dist

    def fn(ptr, n, stream):
        op = dist.ReduceOp.SUM if n > 0 else dist.ReduceOp.MAX
        n = abs(int(n))
        if device is not None and torch.device(device).type == "cuda":
            t = torch.as_tensor(_DevBuf(ptr, n), device=device)
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=device)) if stream else torch.cuda.stream(torch.cuda.current_stream(device)):
                dist.all_reduce(t, op=op, group=group)
        else:   # host memory (CPU tests: gloo over the kernel emulation build)
            arr = np.ctypeslib.as_array((ctypes.c_double * n).from_address(int(ptr)))
            t = torch.from_numpy(arr)
            dist.all_reduce(t, op=op, group=group)
    return fn


class ShardedLBA:
    """One rank of a landmark-sharded LBA.  `solver` is this rank's LBASolver; `P` the FULL problem (every rank holds it,
    or at least its own shard plus the keyframe table)."""

    def __init__(self, solver, rank, world, group=None, device=None, nccl=False):
        self.solver, self.rank, self.world = solver, int(rank), int(world)
        self.shard = self.pt_index = self.ls_index = None
        if nccl:
            import torch
            import torch.distributed as dist
            dev = torch.device(device) if device is not None else torch.device("cpu")
            uid = torch.zeros(128, dtype=torch.uint8, device=dev if dist.get_backend(group) == "nccl" else "cpu")
            if self.rank == 0:
                uid.copy_(torch.tensor(list(self.solver.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, src=0, group=group)
            self.solver.comm_init_rank(self.world, self.rank, bytes(uid.cpu().tolist()))
            self._fn = None
        else:
            self._fn = make_allreduce(group, device)
            self.solver.set_allreduce(self._fn)
            self.solver.set_allreduce_ranks(self.world, self.rank)

    def upload(self, P, opt, masks=None):
        """`masks` = (point mask, line mask) of this rank's landmarks instead of the balanced base-keyframe split (the masks of all ranks
        must partition the landmarks).  An EMPTY shard is legal inside the library — the rank contributes zeros to every exchange and takes the
        same LM decisions —; the default split refuses it because it means the window was spread over too many ranks."""
        if opt.profile != abi.PROFILE_G:
            raise ValueError("sharding is defined for profile G (the Plücker-mode LBA of BASELINE configs 4-5)")
        if masks is not None:
            mp_, ml_ = masks
            self.shard, self.pt_index, self.ls_index = P.subset_landmarks(mp_, ml_), np.flatnonzero(mp_), np.flatnonzero(ml_)
            self.full = P
            self.solver.upload(self.shard, opt)
            return
        # emptiness is decided from the FULL problem, identically on every rank, before anything collective happens: a rank that
        # raised alone would leave the others blocked in the first exchange
        for r in range(self.world):
            mp_, ml_ = shard_masks(P, r, self.world)
            if int(mp_[P.po_lm].sum()) + int(ml_[P.lo_lm].sum()) == 0:
                raise ValueError("rank %d would receive an empty shard: use fewer ranks for this window" % r)
        self.shard, self.pt_index, self.ls_index = shard_problem(P, self.rank, self.world)
        self.full = P
        self.solver.upload(self.shard, opt)

    def run(self):
        self.solver.run()

    def solve(self, P, opt, masks=None):
        """Upload the rank's shard, run the whole LM schedule with the exchange step, return this rank's Result."""
        self.upload(P, opt, masks)
        self.solver.run()
        return self.solver.download()[0]

    def scatter_into(self, res, full_result):
        """Write this rank's landmark outputs into arrays shaped like the full problem (poses are identical on all ranks)."""
        full_result.kf_T_wc[:] = res.kf_T_wc
        full_result.pt_xyz[self.pt_index] = res.pt_xyz
        full_result.ls_orth[self.ls_index] = res.ls_orth
        full_result.ls_plk[self.ls_index] = res.ls_plk
        return full_result
