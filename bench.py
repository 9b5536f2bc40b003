#!/usr/bin/env python
"""bench.py — LBA throughput of the CUDA path on B200, with roofline and CPU baseline (contract: task statement ④).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5] [--profile G|H_END|H_PLK] [--impl reference]

A *step* is one complete local bundle adjustment (the drop-in call: full LM schedule of the chosen reference profile)
over one synthetic window of the named workload (C3: over a batch of 1 024 windows).  Default workload = BASELINE config 2
(KITTI-shaped 20-KF window, 8k points, 2k Plücker lines), default profile = G (what the reference runs in Plücker mode).
  value : observations/s = (point + line observations) x LM trials / device time, problem resident in HBM
          (plba_reset_state + plba_run), CUDA events on the library's stream, L2 flushed between steps.
  e2e   : the same metric through plba_solve() with HOST buffers in and out (H2D + D2H inside the timed region).
  N > 1 : one rank per GPU, one independent window per rank (the same synthetic window: identical per-GPU work), no data-path
          collective: weak scaling.
The same JSON line also carries, measured in the same run (rank 0; --no-configs skips them):
  configs           : whole-LBA resident + end-to-end + 1-thread CPU numbers of the OTHER BASELINE configs (C1, C3, C4, C5),
  roofline_largest  : the assembly kernel at config 5 against the measured HBM peak AND the measured FP64 peak (DFMA / DMMA micro-benchmarks),
  dense_cholesky    : the reduced camera system of config 5 factored as a DENSE 12 000 x 12 000 matrix (the route of loop-closure-shaped windows):
                      n^3 / 3 FP64 flops on mma.sync.m8n8k4.f64 against the DMMA peak measured in the same run (bound: tensor),
  N > 1             : c3_replicas (config 3: 1 024 windows, 1 024 / N per rank, no collective) and sharded (configs 4 and 5, landmarks
                      sharded by base keyframe, the library's own ncclAllReduce per LM trial) next to their single-GPU times.
--impl reference : the CPU restatement of the reference (oracle/, all host threads) on the same workload; rank 0 only; loads neither
                   libplba.so nor CUDA.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PROFILES = {"G": 0, "H_END": 1, "H_PLK": 2}
WORKLOADS = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5}
C3_WINDOWS = 1024
METRIC = "lba_observations_per_s"
UNIT = "observations/s"


def algorithmic_bytes(P, nnzb_S):
    """every input read once + every output written once per LM trial (assembly + Schur stage)."""
    nfix = P.n_kf - P.n_free
    return (32 * P.n_pobs + 48 * P.n_lobs + 24 * P.n_pt + 32 * P.n_ls + 96 * (P.n_free + nfix)
            + 288 * nnzb_S + 48 * P.n_free + 72 * P.n_pt + 112 * P.n_ls)


def algorithmic_flops(P):
    """SURVEY.md §8(d) profile-G estimate, evaluated per landmark track length."""
    kp = np.bincount(P.po_lm, minlength=P.n_pt).astype(np.float64) if P.n_pobs else np.zeros(P.n_pt)
    kl = np.bincount(P.lo_lm, minlength=P.n_ls).astype(np.float64) if P.n_lobs else np.zeros(P.n_ls)
    return float(313.0 * P.n_pobs + 620.0 * P.n_lobs + np.sum(50 + 144 * kp + 108 * kp * (kp + 1)) + np.sum(150 + 240 * kl + 144 * kl * (kl + 1)))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons DURING the timed region (NVML; same fields as the nvidia-smi recipe line)."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.stop_flag = index, period, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload_name(name, P):
    if isinstance(P, (list, tuple)):
        return "%s: %d independent windows of %d free + %d fixed KFs, %d points, %d lines, %d observations in all" % (
            name, len(P), P[0].n_free, P[0].n_kf - P[0].n_free, P[0].n_pt, P[0].n_ls, sum(p.n_obs for p in P))
    return "%s: %d free + %d fixed KFs, %d points, %d lines, %d+%d observations" % (
        name, P.n_free, P.n_kf - P.n_free, P.n_pt, P.n_ls, P.n_pobs, P.n_lobs)


def config_dict(args, P, name=None):
    """The keys are the same in the CUDA arm and in the CPU reference arm."""
    return {"workload": workload_name(name or args.workload, P), "profile": args.profile, "quirks": args.quirks,
            "windows_per_gpu": len(P) if isinstance(P, (list, tuple)) else 1,
            "l2": "flushed between steps (256 MiB write)", "parallelism": "one independent window per GPU (the same synthetic window on every rank), no collective"}


def make_workload(name, prof, scene, abi, windows=None):
    """One Problem (C1, C2, C4, C5) or a list of independent windows (C3: element b uses seed + 1000 b)."""
    lm = 1 if prof == abi.PROFILE_H_END else 0
    if name == "C3":
        seed0 = 20261018 + 3
        idx = range(C3_WINDOWS) if windows is None else windows
        return [scene.make_scene(3, line_mode=lm, seed=seed0 + 1000 * b) for b in idx]
    cfg = WORKLOADS[name]
    return scene.make_scene(cfg, line_mode=lm, seed=int(scene.preset(cfg).seed))


def n_obs_of(P):
    return sum(p.n_obs for p in P) if isinstance(P, (list, tuple)) else P.n_obs


# ------------------------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------------------------
def cut_options(abi, name, prof, quirks):
    """A bounded sample of a config-4 / config-5 window for a CPU: the first outer LM iteration(s) of the same window."""
    if prof == abi.PROFILE_G:
        return abi.Options(prof, quirks, iters_stage1=2 if name == "C4" else 1, iters_stage2=0), "iters_stage1=%d, iters_stage2=0" % (2 if name == "C4" else 1)
    return abi.Options(prof, quirks, max_iters_lba=2), "max_iters_lba=2"


def cpu_sample(orc, abi, name, P, prof, quirks, threads, budget_s):
    """Bounded CPU sample of a workload with the oracle: (observations x trials / s, description)."""
    orc.set_threads(threads)
    if name == "C3":
        sub = P[:8]
        opt = abi.Options(prof, quirks)
        t0 = time.perf_counter(); tr = 0; n = 0
        while n < 1 or time.perf_counter() - t0 < budget_s:
            for p in sub:
                tr += orc.solve(p, opt).n_trials * p.n_obs
            n += 1
        dt = time.perf_counter() - t0
        return tr / dt, "%d passes over the first 8 of the %d windows (full LM schedule each), %.1f s" % (n, len(P), dt)
    if name in ("C4", "C5"):
        opt, how = cut_options(abi, name, prof, quirks)
        t0 = time.perf_counter(); r = orc.solve(P, opt); dt = time.perf_counter() - t0
        return P.n_obs * max(r.n_trials, 1) / dt, "the first %d LM trial(s) of the same window (schedule cut to %s), %.1f s" % (r.n_trials, how, dt)
    opt = abi.Options(prof, quirks)
    t0 = time.perf_counter(); tr = 0; n = 0
    while n < 1 or time.perf_counter() - t0 < budget_s:
        tr += orc.solve(P, opt).n_trials; n += 1
    dt = time.perf_counter() - t0
    return P.n_obs * tr / dt, "%d full LBA solves of the same window, %.1f s" % (n, dt)


def run_reference(args, rank):
    """CPU arm: the oracle (restatement of the reference LBA; the reference application is not buildable here) on all host threads.
    Loads the scene generator (libplba_scene.so, host code) and the oracle only: never libplba.so, never CUDA."""
    if rank != 0:
        return
    from oracle import loader as orc
    from pl_slam_plucker_b200 import abi, scene
    orc.build()
    cores = orc.set_threads(os.cpu_count() or 1)
    prof = PROFILES[args.profile]
    P = make_workload(args.workload, prof, scene, abi)
    opt = abi.Options(prof, args.quirks)

    def one_step():
        if isinstance(P, list):            # C3: a bounded sample of the batch per step (16 windows), OpenMP over windows
            rc, rs = orc.solve_batch(P[:16], opt)
            return sum(r.n_trials * p.n_obs for r, p in zip(rs, P[:16]))
        if args.workload in ("C4", "C5"):  # bounded sample: the first outer iteration(s) of the window
            return P.n_obs * max(orc.solve(P, cut_options(abi, args.workload, prof, args.quirks)[0]).n_trials, 1)
        return P.n_obs * orc.solve(P, opt).n_trials
    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter(); obs_trials = 0
    for _ in range(args.steps):
        obs_trials += one_step()
    dt = time.perf_counter() - t0
    val = obs_trials / dt
    sample = {"C3": "each step = the first 16 of the 1 024 windows, full LM schedule", "C4": "each step = the first outer LM iterations of the window",
              "C5": "each step = the first outer LM iteration of the window"}.get(args.workload, "each step = one full LBA solve of the window")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, P),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%s (oracle/plba_oracle.cpp, OpenMP over landmarks / windows on %d threads; the reference itself is single-threaded, CMakeLists.txt:34: "
                                       "see the CUDA arm's cpu_baseline for the 1-thread number)" % (sample, cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------------
def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


def _traffic(workload):
    """Measured DRAM bytes per assembly-kernel launch from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload]
        return int(t.get("k_assemble_w", t["k_assemble"]) if workload == "C5" else t["k_assemble"])
    except Exception:
        return None


class Timer:
    """Resident and end-to-end whole-LBA timing of one workload on one handle."""

    def __init__(self, torch, s, stream, flush, barrier):
        self.torch, self.s, self.stream, self.flush, self.barrier = torch, s, stream, flush, barrier

    def resident(self, P, opt, steps, warmup):
        torch, s, stream = self.torch, self.s, self.stream
        s.upload(P, opt)
        for _ in range(warmup):
            s.reset(); s.run()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        trials = launches = 0
        self.barrier()
        for k in range(steps):
            with torch.cuda.stream(stream):
                # L2 flush between timed iterations (untimed).  Written twice: the second write is still running when the host enqueues
                # the start event and the LBA, so the timed region never contains the host's launch latency
                self.flush.zero_(); self.flush.zero_()
                s.reset()
                ev[k][0].record(stream)
                s.run()
                ev[k][1].record(stream)
            t = s.timing()
            launches += t["n_launches_run"]; trials += t["n_trials_run"]
        self.barrier()
        return sum(a.elapsed_time(b) for a, b in ev), trials, launches

    def stages(self, steps):
        """The same steps replayed by the host-driven loop with CUDA events around each stage (resident upload)."""
        torch, s, stream = self.torch, self.s, self.stream
        s.set_detail_timing(True)
        t_asm = t_sol = t_upd = 0.0; n_asm = 0
        for _ in range(steps):
            with torch.cuda.stream(stream):
                self.flush.zero_(); s.reset(); s.run()
            t = s.timing()
            t_asm += t["ms_assemble"]; t_sol += t["ms_solve"]; t_upd += t["ms_update"]; n_asm += t["n_assemble_run"]
        s.set_detail_timing(False)
        return t_asm, t_sol, t_upd, n_asm

    def e2e(self, P, opt, steps, abi):
        s = self.s
        batch = isinstance(P, (list, tuple))
        if batch:
            bufs = s.batch_buffers(P, 64)
            call = lambda: s.solve_batch(P, opt, out=bufs)[1]      # noqa: E731
        else:
            out = abi.Result(P, 256)                   # caller-owned result buffers, re-used every call like a SLAM back-end would
            call = lambda: [s.solve(P, opt, out=out)]  # noqa: E731
        call()
        self.barrier()
        t0 = time.perf_counter(); tr = 0; h2d = d2h = 0; hp = hu = gpu_ms = 0.0
        for _ in range(steps):
            rs = call()
            tr += sum(r.n_trials * p.n_obs for r, p in zip(rs, P if batch else [P]))
            t = s.timing(); h2d += t["h2d_bytes"]; d2h += t["d2h_bytes"]; hp += t["ms_host_prep"]; hu += t["ms_host_unpack"]; gpu_ms += t["ms_total"]
        self.barrier()
        dt = time.perf_counter() - t0
        return dt, tr, h2d // steps, d2h // steps, {"host_flatten": hp / steps, "device_lm_loop": gpu_ms / steps, "host_unpack": hu / steps}


def fp64_peaks(s):
    try:
        return {"dfma_tflops": s.measure_fp64_peak(0, 3), "dmma_tflops": s.measure_fp64_peak(1, 3),
                "how": "k_peak_dfma / k_peak_dmma in libplba.so: independent FMA chains / mma.sync.m8n8k4.f64, 8 x SMs CTAs of 256 threads, best of 3 launches, CUDA events"}
    except Exception as e:      # noqa: BLE001
        return {"error": str(e)}


def largest_config_roofline(s, abi, args, P5, peaks):
    """Assembly kernel on BASELINE config 5 (2 000 KFs, 2M points, 500k lines; inputs 0.45 GB >> L2, so every launch is cold)."""
    s.upload(P5, abi.Options(PROFILES[args.profile] if args.profile != "H_END" else 0, args.quirks))
    st = s.layout_stats(); kp = s.kernel_path()
    A = algorithmic_bytes(P5, st["nnzb_S"]); F = algorithmic_flops(P5)
    ms_a = s.time_kernel(0, 5); ms_u = s.time_kernel(2, 5); ms_s = s.time_kernel(1, 3)
    peak, src = _peak()
    fp = peaks.get("dfma_tflops")
    return {"workload": "C5: %d free KFs, %d points, %d lines, %d+%d observations" % (P5.n_free, P5.n_pt, P5.n_ls, P5.n_pobs, P5.n_lobs),
            "kernel": "k_assemble_w" if kp["assembly"] == "warp" else "k_assemble", "kernel_path": kp, "bound": "hbm", "algorithmic_bytes": A, "ms_per_launch": ms_a, "achieved": A / (ms_a * 1e-3) / 1e9, "peak": peak,
            "peak_source": src, "unit": "GB/s", "frac": A / (ms_a * 1e-3) / 1e9 / peak, "traffic": _traffic("C5"), "algorithmic_flops": F, "achieved_fp64_tflops": F / (ms_a * 1e-3) / 1e12,
            "peak_fp64_measured_tflops": fp, "frac_fp64": (F / (ms_a * 1e-3) / 1e12 / fp) if fp else None,
            "update_kernel_ms_per_launch": ms_u, "solve_ms_per_trial": ms_s, "lm_trial_ms": ms_a + ms_u + ms_s,
            "observations_per_s_per_trial": P5.n_obs / ((ms_a + ms_u + ms_s) * 1e-3), "nnzb_S": st["nnzb_S"], "launches_timed": 5,
            "note": "FP64 work (%.1f algorithmic GFLOP) bounds this kernel before HBM does: at the measured DFMA peak it needs %.0f us, at the HBM peak %.0f us; "
                    "frac is against the HBM roof (BASELINE's metric), frac_fp64 against the measured FP64 roof" % (F / 1e9, F / ((fp or 37.0) * 1e12) * 1e6, A / (peak * 1e9) * 1e6)}


def dense_cholesky_roofline(s, abi, args, P5, peaks):
    """The reduced camera system of config 5 (12 000 unknowns) factored as a DENSE matrix — the route a loop-closure-shaped window of that
    size takes (its reduced system is not block-banded, so the block cyclic reduction does not apply).  Tensor-bound: n^3 / 3 FP64 flops on
    mma.sync.m8n8k4.f64 against the DMMA peak measured in the same run."""
    try:
        s.set_force_dense(True)
        s.upload(P5, abi.Options(PROFILES[args.profile] if args.profile != "H_END" else 0, args.quirks))
        n = 6 * P5.n_free
        ms = s.time_kernel(1, 3)
        fl = n ** 3 / 3.0
        fp = peaks.get("dmma_tflops")
        out = {"workload": "dense reduced camera system of C5: n = %d (factorisation + both substitutions)" % n, "kernel": "k_syrk_dmma (+ k_potrf_block, k_trsm_block, k_rhs_update, k_back_block, k_back_update)",
               "bound": "tensor", "ms_per_solve": ms, "flops": fl, "achieved": fl / (ms * 1e-3) / 1e12, "peak": fp, "unit": "TFLOP/s", "frac": (fl / (ms * 1e-3) / 1e12 / fp) if fp else None,
               "peak_source": "measured in this run (k_peak_dmma)", "kernel_path": s.kernel_path(), "solves_timed": 3}
        s.set_force_dense(False)
        try:
            # the same size with loop closures (keyframe i also observes landmarks of keyframe i - 500): the reduced camera system is not banded,
            # plba_upload routes to the dense solver by itself; whole LBA, schedule cut to 2 + 1 outer iterations to bound the run
            from pl_slam_plucker_b200 import scene as _scene
            PL = _scene.make_scene(5, loop_every=500, seed=int(_scene.preset(5).seed))
            optL = abi.Options(abi.PROFILE_G, args.quirks, iters_stage1=2, iters_stage2=1)
            s.upload(PL, optL)
            best = None
            for _ in range(2):
                s.reset(); t0 = time.perf_counter(); s.run(); dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            tmg = s.timing()
            trials = max(int(tmg.get("n_trials_run", 0)), 1)
            stage = None
            try:
                s.set_detail_timing(True); s.reset(); s.run(); td = s.timing()
                stage = {k: td[k] for k in ("ms_assemble", "ms_solve", "ms_update", "ms_other") if k in td}
            finally:
                s.set_detail_timing(False)
            out["loop_closure_window"] = {"workload": "C5 with loop closures every 500 keyframes: %d free KFs, %d points, %d lines, %d observations (2 + 1 outer iterations)" % (PL.n_free, PL.n_pt, PL.n_ls, PL.n_obs),
                                          "kernel_path": s.kernel_path(), "ms_per_lba": 1e3 * best, "lm_trials": trials, "ms_per_trial": 1e3 * best / trials, "stage_ms_per_lba": stage,
                                          "how": "plba_reset_state + plba_run, wall clock around the synchronous call, best of 2"}
        except Exception as e:      # noqa: BLE001
            out["loop_closure_window"] = {"error": str(e)}
        return out
    except Exception as e:      # noqa: BLE001
        return {"error": str(e)}
    finally:
        s.set_force_dense(False)


def other_configs(tm, s, abi, scene, orc, args, prof, p5_holder):
    """Whole-LBA numbers of the BASELINE configs other than the headline one: resident, end-to-end, 1-thread CPU sample."""
    out = {}
    for name in ("C1", "C3", "C4", "C5"):
        if name == args.workload:
            continue
        P = make_workload(name, prof, scene, abi)
        if name == "C5":
            p5_holder.append(P)
        opt = abi.Options(prof, args.quirks)
        steps = {"C1": 10, "C3": 3, "C4": 5, "C5": 2}[name]
        e_steps = {"C1": 5, "C3": 2, "C4": 2, "C5": 1}[name]
        ms, trials, launches = tm.resident(P, opt, steps, 2)
        kp = s.kernel_path()
        nobs = n_obs_of(P)
        ent = {"workload": workload_name(name, P), "ms_per_lba": ms / steps, "lm_trials_per_lba": trials / steps,
               "observations_per_s": nobs * trials / steps / (ms / steps * 1e-3), "kernel_path": kp, "gpu_launches_per_lba": launches / steps}
        if isinstance(P, list):      # per-window trial counts differ: observations x trials summed over windows
            rs = s.download(64)
            ent["observations_per_s"] = sum(r.n_trials * p.n_obs for r, p in zip(rs, P)) / (ms / steps * 1e-3)
            ent["windows_per_s"] = len(P) / (ms / steps * 1e-3)
        dt, e_tr, h2d, d2h, brk = tm.e2e(P, opt, e_steps, abi)
        ent["e2e"] = {"value": e_tr / dt, "unit": UNIT, "ms_per_lba": 1e3 * dt / e_steps, "h2d_bytes": int(h2d), "d2h_bytes": int(d2h), "breakdown_ms": brk}
        if orc is not None:
            v, desc = cpu_sample(orc, abi, name, P, prof, args.quirks, 1, 3.0)
            ent["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc}
        out[name] = ent
    return out


def c3_replicas(tm, s, abi, scene, args, prof, rank, world, torch, dist, dev):
    """BASELINE config 3 as it is worded: 1 024 independent windows, 1 024 / N per GPU, no collective."""
    mine = list(range(rank, C3_WINDOWS, world))
    P = make_workload("C3", prof, scene, abi, windows=mine)
    opt = abi.Options(prof, args.quirks)
    steps = 3
    ms, trials, launches = tm.resident(P, opt, steps, 2)
    rs = s.download(64)
    obs_trials = float(sum(r.n_trials * p.n_obs for r, p in zip(rs, P)))
    t = torch.tensor([ms / steps], dtype=torch.float64, device=dev); c = torch.tensor([obs_trials, float(len(P))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return {"workload": "C3: %d independent 10-KF windows, %d per GPU, no data-path collective" % (int(c[1].item()), len(P)), "ms_per_batch_lba": float(t.item()),
            "observations_per_s": float(c[0].item()) / (float(t.item()) * 1e-3), "windows_per_s": float(c[1].item()) / (float(t.item()) * 1e-3), "scaling": "strong (1 024 windows in all)"}


def sharded_run(abi, scene, args, rank, world, dev, stream, barrier, torch, dist, name):
    """A BASELINE config-4 / config-5 window sharded by base keyframe over the ranks; the library itself all-reduces the reduced camera
    system with NCCL once per LM trial (plba_comm_init_rank).  Next to it: the same window on ONE GPU (rank 0), same run."""
    from pl_slam_plucker_b200 import sharded, solver
    P = make_workload(name, abi.PROFILE_G, scene, abi)
    opt = abi.Options(abi.PROFILE_G, args.quirks)
    s2 = solver.LBASolver(dev.index, stream=stream.cuda_stream)
    sh = sharded.ShardedLBA(s2, rank, world, device=dev, nccl=True)
    reps = 3 if name == "C4" else 2
    with torch.cuda.stream(stream):
        sh.upload(P, opt)
        s2.run()                                   # warm-up (NCCL channels, allocator)
        t_ms, trials = 0.0, 0
        for _ in range(reps):
            s2.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier(); e0.record(stream); s2.run(); e1.record(stream); torch.cuda.synchronize()
            t_ms += e0.elapsed_time(e1); trials += s2.timing()["n_trials_run"]
    t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_dbl = s2.reduced_system_ptr()[1]; kpath = s2.kernel_path()
    s2.comm_destroy(); s2.close()
    single = None
    if rank == 0:                                  # the strong-scaling denominator, measured in the same run on one of the same GPUs
        s1 = solver.LBASolver(dev.index, stream=stream.cuda_stream)
        with torch.cuda.stream(stream):
            s1.upload(P, opt); s1.run(); ts = 0.0
            for _ in range(reps):
                s1.reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); s1.run(); e1.record(stream); torch.cuda.synchronize()
                ts += e0.elapsed_time(e1)
        single = ts / reps
        s1.close()
    barrier()
    ms = float(t.item()) / reps
    return {"workload": "%s: %d free KFs, %d points, %d lines sharded by base keyframe over %d GPUs" % (name, P.n_free, P.n_pt, P.n_ls, world),
            "lm_trials": trials // reps, "ms_per_lba": ms, "observations_per_s": P.n_obs * (trials / reps) / (ms * 1e-3),
            "single_gpu_ms_per_lba": single, "speedup_vs_single_gpu": (single / ms) if single else None,
            "allreduce_bytes_per_trial": int(8 * n_dbl), "kernel_path": kpath,
            "collective": "ncclAllReduce issued by the library on its own stream (plba_comm_init_rank): one per LM trial (the band of the reduced camera system in node form "
                          "[D | U | b | cost sums] for the block-cyclic-reduction solver, the dense [S | g] otherwise) + 4 doubles after the update kernel; no host "
                          "synchronisation inside the LM loop; the reduced-system solve is replicated on every rank"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--profile", default="G", choices=sorted(PROFILES))
    ap.add_argument("--quirks", type=int, default=0, help="0 = faithful (bug-for-bug), 1 = fixed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-largest", action="store_true", help="skip the C5 assembly roofline measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the whole-LBA numbers of the other BASELINE configs")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the config-3 replica run and the landmark-sharded runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from pl_slam_plucker_b200 import abi, scene, solver
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the LBA path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    prof = PROFILES[args.profile]
    # weak scaling: every rank solves its own copy of the SAME synthetic window, so that the per-GPU work is identical (windows drawn with
    # different seeds take 19-25 LM trials and the step time, a maximum over ranks, would measure that imbalance instead)
    P = make_workload(args.workload, prof, scene, abi)
    opt = abi.Options(prof, args.quirks)
    stream = torch.cuda.Stream(device=dev)
    s = solver.LBASolver(local, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tm = Timer(torch, s, stream, flush, barrier)
    sampler = ClockSampler(local); sampler.start()
    # ---- resident-problem throughput: the whole LM schedule is one CUDA-graph launch per step ----
    ms, trials, launches = tm.resident(P, opt, args.steps, args.warmup)
    nnzb = s.layout_stats()["nnzb_S"]; kpath = s.kernel_path()
    if isinstance(P, list):
        obs_trials_rank = float(sum(r.n_trials * p.n_obs for r, p in zip(s.download(64), P))) * args.steps
    else:
        obs_trials_rank = float(P.n_obs * trials)
    # ---- per-kernel durations: the same steps replayed by the host-driven loop with CUDA events around each stage ----
    det_steps = max(3, min(args.steps, 20))
    t_asm, t_sol, t_upd, n_asm = tm.stages(det_steps)
    # ---- end to end through the drop-in call: host buffers in, host buffers out ----
    e2e_steps = max(3, args.steps // 4)
    e2e_s, e2e_obs_trials, h2d, d2h, brk = tm.e2e(P, opt, e2e_steps, abi)
    sampler.stop_flag = True; sampler.join()

    extras = {}
    orc = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import loader as orc
        orc.build()
    p5 = []
    if rank == 0 and world == 1 and not args.no_configs:
        extras["configs"] = other_configs(tm, s, abi, scene, orc, args, prof, p5)
    if rank == 0 and not args.no_largest:
        peaks = fp64_peaks(s)
        extras["fp64_peaks"] = peaks
        extras["roofline_largest"] = largest_config_roofline(s, abi, args, p5[0] if p5 else make_workload("C5", 0, scene, abi), peaks)
        extras["dense_cholesky"] = dense_cholesky_roofline(s, abi, args, p5[0] if p5 else make_workload("C5", 0, scene, abi), peaks)
    if world > 1 and not args.no_sharded:
        extras["c3_replicas"] = c3_replicas(tm, s, abi, scene, args, prof, rank, world, torch, dist, dev)
        extras["sharded"] = {name: sharded_run(abi, scene, args, rank, world, dev, stream, barrier, torch, dist, name) for name in ("C4", "C5")}
    tt = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    cnt = torch.tensor([obs_trials_rank, float(e2e_obs_trials), float(trials), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = tt.tolist(); obs_trials, e2e_obs_trials_all, trials_all, launches_all = cnt.tolist()

    if rank == 0:
        peak, peak_src = _peak()
        P1 = P[0] if isinstance(P, list) else P
        A = (sum(algorithmic_bytes(p, 0) for p in P) + 288 * nnzb) if isinstance(P, list) else algorithmic_bytes(P, nnzb)
        asm_ms = t_asm / max(n_asm, 1)
        line = {"metric": METRIC, "value": obs_trials / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "lm_iters_per_s": trials_all / (ms_max * 1e-3), "lm_trials_per_step": trials / args.steps,
                "config": config_dict(args, P),
                "e2e": {"value": e2e_obs_trials_all / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_ms_max / e2e_steps, "steps": e2e_steps, "breakdown_ms": brk},
                "gpu_launches": int(launches_all),
                "gpu_launches_how": "counted on the device: every kernel of the library increments a counter when it starts (the LM loop is one CUDA graph per step)",
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "kernel": ("k_assemble_w" if kpath["assembly"] == "warp" else "k_assemble") + " (one launch per LM trial: points + lines)", "kernel_path": kpath, "achieved": A / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else None,
                             "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": (A / (asm_ms * 1e-3) / 1e9 / peak) if asm_ms > 0 else None,
                             "traffic": _traffic(args.workload) if args.profile == "G" else None, "algorithmic_bytes": A, "ms_per_launch": asm_ms, "launches_timed": int(n_asm),
                             "stage_ms_per_step": {"assemble": t_asm / det_steps, "solve": t_sol / det_steps, "update": t_upd / det_steps},
                             "how": "CUDA events around each stage in a host-driven replay of the same steps (the headline steps run as one CUDA graph)",
                             "note": "BASELINE's metric quotes the roofline on the assembly stage, so this object is the assembly kernel; a %d-unknown window is latency-bound in "
                                     "every kernel (reduced-system Cholesky %.0f %% of the stage time): see roofline_largest for the configuration whose assembly is "
                                     "throughput-bound" % (6 * P1.n_free, 100.0 * t_sol / max(t_asm + t_sol + t_upd, 1e-9))}}
        line.update(extras)
        if orc is not None:
            v, desc = cpu_sample(orc, abi, args.workload, P, prof, args.quirks, 1, 10.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": desc + " (oracle, single thread like the reference: CMakeLists.txt:34 has no OpenMP)", "host_cores_available": os.cpu_count()}
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
