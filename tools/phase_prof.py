"""Per-phase cycle breakdown of the stage kernels (needs the -DPLBA_PROF build): python tools/phase_prof.py <cfg>..."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver, _lib
lib = _lib.load(os.path.join(os.path.dirname(_lib.library_path()), "libplba_prof.so"))
s = solver.LBASolver(0, lib=lib)
NAMES = {1: "asm precompute", 2: "asm linearize", 3: "asm per-landmark", 4: "asm Ta", 5: "asm offdiag tasks", 6: "asm diag tasks",
         21: "upd precompute", 27: "upd linearize (old state: before the solver's flag)", 22: "upd J x_p", 23: "upd per-landmark", 24: "upd orth->plk", 25: "upd new cost", 26: "upd tail",
         42: "sol control", 43: "sol load + first factor", 44: "sol A: column solve", 45: "sol B: trailing update + look-ahead factor", 46: "sol (loop exit)", 47: "sol backward", 50: "sol write + pose"}
for cfg in [int(a) for a in sys.argv[1:]] or [2]:
    P = scene.make_scene(cfg)
    s.upload(P, abi.Options(abi.PROFILE_G, 1))
    for which in (0, 1, 2):
        if which == 1 and P.n_free > 24: continue
        buf = (C.c_ulonglong * 64)()
        lib.plba_debug_prof(buf, 1)
        reps = 5
        ms = s.time_kernel(which, reps)
        lib.plba_debug_prof(buf, 0)
        tot = sum(buf)
        print("cfg %d kernel %d: %.1f us/launch (instrumented); CTA-cycles by phase:" % (cfg, which, 1e3 * ms))
        for i in range(64):
            if buf[i]: print("   %-24s %6.1f%%  %10.0f cycles/launch" % (NAMES.get(i, str(i)), 100.0 * buf[i] / tot, buf[i] / (reps + 2)))
