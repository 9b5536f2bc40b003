"""ctypes binding of libplba.so (the C ABI of include/plba.h).  CUDA-only: there is no CPU fallback."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_pd = C.POINTER(C.c_double)
ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)

EXPORTS = ["plba_version", "plba_default_options", "plba_create", "plba_destroy", "plba_last_error", "plba_solve", "plba_solve_batch",
           "plba_upload", "plba_reset_state", "plba_run", "plba_download", "plba_trial_assemble", "plba_trial_finish",
           "plba_reduced_system", "plba_copy_reduced_system", "plba_time_kernel", "plba_layout_stats", "plba_kernel_path", "plba_track_default_options", "plba_track_solve", "plba_create_lines", "plba_set_force_dense", "plba_set_force_chunk", "plba_set_allreduce", "plba_set_allreduce_ranks", "plba_comm_unique_id", "plba_comm_init_rank", "plba_comm_destroy", "plba_create_group", "plba_destroy_group", "plba_comm_info", "plba_measure_fp64_peak", "plba_get_timing", "plba_set_detail_timing", "plba_scene_preset", "plba_scene_create",
           "plba_scene_problem", "plba_scene_truth", "plba_scene_destroy"]


def declare(L):
    abi.declare_common(L)
    L.plba_version.restype = C.c_int
    L.plba_default_options.argtypes = [C.c_int32, C.POINTER(abi.plba_options)]
    L.plba_default_options.restype = None
    L.plba_create.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
    L.plba_create.restype = C.c_int
    L.plba_destroy.argtypes = [C.c_void_p]
    L.plba_destroy.restype = None
    L.plba_last_error.argtypes = [C.c_void_p]
    L.plba_last_error.restype = C.c_char_p
    L.plba_solve.argtypes = [C.c_void_p, C.POINTER(abi.plba_problem), C.POINTER(abi.plba_options), C.POINTER(abi.plba_result)]
    L.plba_solve.restype = C.c_int
    L.plba_solve_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(abi.plba_problem), C.POINTER(abi.plba_options), C.POINTER(abi.plba_result)]
    L.plba_solve_batch.restype = C.c_int
    L.plba_upload.argtypes = [C.c_void_p, C.c_int32, C.POINTER(abi.plba_problem), C.POINTER(abi.plba_options)]
    L.plba_upload.restype = C.c_int
    L.plba_reset_state.argtypes = [C.c_void_p]
    L.plba_reset_state.restype = C.c_int
    L.plba_run.argtypes = [C.c_void_p]
    L.plba_run.restype = C.c_int
    L.plba_download.argtypes = [C.c_void_p, C.c_int32, C.POINTER(abi.plba_result)]
    L.plba_download.restype = C.c_int
    L.plba_trial_assemble.argtypes = [C.c_void_p, C.c_double]
    L.plba_trial_assemble.restype = C.c_int
    L.plba_trial_finish.argtypes = [C.c_void_p, C.c_double, _pd, _pd]
    L.plba_trial_finish.restype = C.c_int
    L.plba_reduced_system.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.plba_reduced_system.restype = C.c_int
    L.plba_copy_reduced_system.argtypes = [C.c_void_p, C.c_int32, _pd, _pd]
    L.plba_copy_reduced_system.restype = C.c_int
    L.plba_time_kernel.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, _pd]
    L.plba_time_kernel.restype = C.c_int
    L.plba_set_force_dense.argtypes = [C.c_void_p, C.c_int]
    L.plba_set_force_dense.restype = C.c_int
    L.plba_kernel_path.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.plba_kernel_path.restype = C.c_int
    L.plba_set_force_chunk.argtypes = [C.c_void_p, C.c_int]
    L.plba_set_force_chunk.restype = C.c_int
    L.plba_layout_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.plba_layout_stats.restype = C.c_int
    L.plba_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p]
    L.plba_set_allreduce.restype = C.c_int
    L.plba_set_allreduce_ranks.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.plba_set_allreduce_ranks.restype = C.c_int
    L.plba_comm_unique_id.argtypes = [C.c_void_p]
    L.plba_comm_unique_id.restype = C.c_int
    L.plba_comm_init_rank.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.plba_comm_init_rank.restype = C.c_int
    L.plba_comm_destroy.argtypes = [C.c_void_p]
    L.plba_comm_destroy.restype = C.c_int
    L.plba_create_group.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
    L.plba_create_group.restype = C.c_int
    L.plba_destroy_group.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    L.plba_destroy_group.restype = None
    L.plba_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.plba_comm_info.restype = C.c_int
    L.plba_measure_fp64_peak.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _pd]
    L.plba_measure_fp64_peak.restype = C.c_int
    L.plba_get_timing.argtypes = [C.c_void_p, C.POINTER(abi.plba_timing)]
    L.plba_get_timing.restype = C.c_int
    L.plba_set_detail_timing.argtypes = [C.c_void_p, C.c_int]
    L.plba_set_detail_timing.restype = C.c_int
    return L


_SCENE = None


def load_scene():
    """The synthetic-scene generator (libplba_scene.so: host code only, no CUDA).  Scenes never come from libplba.so, so that the CPU
    reference arm of bench.py and the oracle tests do not map the product library."""
    global _SCENE
    if _SCENE is None:
        so = os.path.join(_HERE, "libplba_scene.so")
        if not os.path.exists(so):
            from . import build as _b
            _b.build_scene()
        L = C.CDLL(so)
        abi.declare_common(L)
        _SCENE = L
    return _SCENE


def library_path():
    return os.path.join(_HERE, "libplba.so")


def load(path=None):
    """Load the CUDA library.  Raises (never falls back) if it has not been built."""
    global _LIB
    if path is not None:
        return declare(C.CDLL(path))
    if _LIB is None:
        so = library_path()
        if not os.path.exists(so):
            raise RuntimeError("%s is missing: run `python -m pl_slam_plucker_b200.build` (nvcc, sm_100a). "
                               "The LBA path is CUDA-only; there is no CPU fallback." % so)
        _LIB = declare(C.CDLL(so))
    return _LIB
