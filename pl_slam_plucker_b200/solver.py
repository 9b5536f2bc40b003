"""Python face of the C ABI: one `LBASolver` per GPU / stream (mirrors plba_handle)."""
import ctypes as C
import numpy as np

from . import abi, _lib


class LBAError(RuntimeError):
    pass


class LBASolver:
    """B200 LBA solver handle.  `solve()` is the drop-in for the reference's three LBA functions (see include/plba.h)."""

    def __init__(self, device=0, stream=None, lib=None, _handle=None):
        self.L = lib or _lib.load()
        self.h = C.c_void_p()
        self._owned = _handle is None
        if _handle is not None:
            self.h = C.c_void_p(_handle)                 # a member of a plba_create_group() group (LBAGroup owns it)
        else:
            rc = self.L.plba_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self.h))
            if rc != 0 or not self.h:
                raise LBAError("plba_create(device=%d) failed (%d): a CUDA device is required; there is no CPU fallback" % (device, rc))
        self._cb = None
        self._probs = None

    def close(self):
        if self.h and self._owned:
            self.L.plba_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=(abi.OK, abi.DISCARDED)):
        if rc not in allow:
            raise LBAError("plba error %d: %s" % (rc, self.L.plba_last_error(self.h).decode()))
        return rc

    # ---- one-call interface (host buffers in, host buffers out) ----
    def solve(self, prob, opt, trace_cap=256, out=None):
        """The drop-in call.  `out`: a Result of matching shape to write into (a steady-state caller re-uses its buffers)."""
        res = out if out is not None else abi.Result(prob, trace_cap)
        pc = prob.as_c()
        res.rc = self._check(self.L.plba_solve(self.h, C.byref(pc), C.byref(opt.c), C.byref(res.c)))
        return res

    def batch_buffers(self, probs, trace_cap=64):
        """Caller-owned argument / result arrays of a batch call, built once and re-used (a steady-state caller does not re-allocate)."""
        n = len(probs)
        arr = (abi.plba_problem * n)(*[p.as_c() for p in probs])
        results = [abi.Result(p, trace_cap) for p in probs]
        rarr = (abi.plba_result * n)(*[r.c for r in results])
        return arr, results, rarr

    def solve_batch(self, probs, opt, trace_cap=64, out=None):
        n = len(probs)
        arr, results, rarr = out if out is not None else self.batch_buffers(probs, trace_cap)
        rc = self._check(self.L.plba_solve_batch(self.h, n, arr, C.byref(opt.c), rarr))
        for i, r in enumerate(results):
            r.c = rarr[i]
            r.rc = rarr[i].status
        return rc, results

    # ---- staged interface (problem resident in HBM) ----
    def upload(self, probs, opt):
        if isinstance(probs, abi.Problem):
            probs = [probs]
        self._probs = list(probs)
        n = len(self._probs)
        arr = (abi.plba_problem * n)(*[p.as_c() for p in self._probs])
        self._check(self.L.plba_upload(self.h, n, arr, C.byref(opt.c)))

    def reset(self):
        self._check(self.L.plba_reset_state(self.h))

    def run(self):
        self._check(self.L.plba_run(self.h))

    def download(self, trace_cap=256):
        n = len(self._probs)
        results = [abi.Result(p, trace_cap) for p in self._probs]
        rarr = (abi.plba_result * n)(*[r.c for r in results])
        self._check(self.L.plba_download(self.h, n, rarr))
        for i, r in enumerate(results):
            r.c = rarr[i]
            r.rc = rarr[i].status
        return results

    def trial_assemble(self, lam):
        self._check(self.L.plba_trial_assemble(self.h, float(lam)))

    def trial_finish(self, lam):
        chi, sc = C.c_double(0), C.c_double(0)
        self._check(self.L.plba_trial_finish(self.h, float(lam), C.byref(chi), C.byref(sc)))
        return chi.value, sc.value

    def reduced_system_ptr(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.L.plba_reduced_system(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def copy_reduced_system(self, window=0):
        """(S, g) of one window after trial_assemble(): S dense [6nf, 6nf] with the upper block triangle filled, undamped."""
        n = 6 * self._probs[window].n_free
        S, g = np.zeros((n, n)), np.zeros(n)
        self._check(self.L.plba_copy_reduced_system(self.h, int(window), S.ctypes.data_as(_lib._pd), g.ctypes.data_as(_lib._pd)))
        return S, g

    def time_kernel(self, which, reps=20, lam=1.0):
        """Average ms of one launch of a stage kernel (0 assemble, 1 solve, 2 update) on the resident problem."""
        ms = C.c_double(0)
        self._check(self.L.plba_time_kernel(self.h, int(which), int(reps), float(lam), C.byref(ms)))
        return ms.value

    def measure_fp64_peak(self, which=0, reps=3):
        """Measured FP64 roof in TFLOP/s: which 0 = DFMA (vector pipe), 1 = DMMA (mma.sync.m8n8k4.f64)."""
        v = C.c_double(0)
        self._check(self.L.plba_measure_fp64_peak(self.h, int(which), int(reps), C.byref(v)))
        return v.value

    def set_force_dense(self, on=True):
        """Large windows: always use the dense DMMA Cholesky instead of the banded one."""
        self._check(self.L.plba_set_force_dense(self.h, 1 if on else 0))

    def set_kernel_path(self, mode=0):
        """Assembly / update kernels: 0 = route by size, 1 = CTA-chunk kernels, 2 = warp-autonomous kernels whenever tracks fit a warp."""
        self._check(self.L.plba_set_force_chunk(self.h, int(mode)))

    def layout_stats(self):
        out = (C.c_int64 * 8)()
        self._check(self.L.plba_layout_stats(self.h, out))
        keys = ("chunks_pt", "chunks_ls", "segments_pt", "segments_ls", "tasks_offdiag", "tasks_diag", "nnzb_S", "arena_bytes")
        return dict(zip(keys, [int(v) for v in out]))

    def kernel_path(self):
        """Kernels the resident upload runs on (plba_kernel_path)."""
        out = (C.c_int32 * 4)()
        self._check(self.L.plba_kernel_path(self.h, out))
        return {"assembly": ("cta-chunk", "warp")[out[0]], "solver": ("shared-memory", "block-cyclic-reduction", "banded", "dense-dmma")[out[1]],
                "band_blocks": int(out[2]), "work_units": int(out[3])}

    def set_allreduce(self, fn):
        """fn(ptr:int, n_doubles:int (negative => max-reduce of |n|), stream:int) sums the device buffer over ranks in place."""
        def _cb(ptr, n, stream, user):
            fn(ptr, n, stream)
        self._cb = _lib.ALLREDUCE_FN(_cb)
        self._check(self.L.plba_set_allreduce(self.h, self._cb, None))

    # ---- landmark-sharded multi-GPU: the collective lives in the library (NCCL on the handle's stream) ----
    def comm_unique_id(self):
        """128 bytes that rank 0 hands to the other ranks (any transport) before comm_init_rank()."""
        buf = (C.c_ubyte * 128)()
        self._check(self.L.plba_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, nranks, rank, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.L.plba_comm_init_rank(self.h, int(nranks), int(rank), buf))

    def comm_destroy(self):
        self._check(self.L.plba_comm_destroy(self.h))

    def comm_info(self):
        out = (C.c_int32 * 2)()
        self._check(self.L.plba_comm_info(self.h, out))
        return int(out[0]), int(out[1])

    def set_allreduce_ranks(self, nranks, rank):
        self._check(self.L.plba_set_allreduce_ranks(self.h, int(nranks), int(rank)))

    def set_detail_timing(self, on=True):
        self.L.plba_set_detail_timing(self.h, 1 if on else 0)

    def timing(self):
        t = abi.plba_timing()
        self.L.plba_get_timing(self.h, C.byref(t))
        return {k: getattr(t, k) for k, _ in abi.plba_timing._fields_}


class LBAGroup:
    """One process, several GPUs: plba_create_group() = one handle per device + one NCCL communicator over them (ncclCommInitAll).
    Drive each member from its own host thread (every member takes part in every collective)."""

    def __init__(self, devices, lib=None):
        self.L = lib or _lib.load()
        n = len(devices)
        devs = (C.c_int32 * n)(*[int(d) for d in devices])
        self._hs = (C.c_void_p * n)()
        rc = self.L.plba_create_group(n, devs, self._hs)
        if rc != 0:
            raise LBAError("plba_create_group(%s) failed (%d)" % (list(devices), rc))
        self.members = [LBASolver(lib=self.L, _handle=self._hs[i]) for i in range(n)]

    def close(self):
        if self._hs is not None:
            self.L.plba_destroy_group(len(self.members), self._hs)
            for m in self.members:
                m.h = C.c_void_p()
            self._hs = None
