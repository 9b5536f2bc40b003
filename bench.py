#!/usr/bin/env python
"""bench.py — LBA throughput of the CUDA path on B200, with roofline and CPU baseline (contract: task statement ④).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C4|C5] [--profile G|H_END|H_PLK] [--impl reference]

A *step* is one complete local bundle adjustment (the drop-in call: full LM schedule of the chosen reference profile)
over one synthetic window of the named workload.  Default workload = BASELINE config 2 (KITTI-shaped 20-KF window,
8k points, 2k Plücker lines), default profile = G (what the reference runs in Plücker mode, SURVEY.md "Read this first").
  value : observations/s = (point + line observations) x LM trials / device time, problem resident in HBM
          (plba_reset_state + plba_run), CUDA events on the library's stream, L2 flushed between steps.
  e2e   : the same metric through plba_solve() with HOST buffers in and out (H2D + D2H inside the timed region).
  N > 1 : one rank per GPU, one independent window per rank (the same synthetic window: identical per-GPU work), no data-path
          collective: weak scaling.
--impl reference : the CPU restatement of the reference (oracle/, all host threads) on the same workload; rank 0 only.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PROFILES = {"G": 0, "H_END": 1, "H_PLK": 2}
WORKLOADS = {"C1": 1, "C2": 2, "C4": 4, "C5": 5}
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


def run_reference(args, rank):
    """CPU arm: the oracle (restatement of the reference LBA; the reference itself is not buildable here) on all host threads."""
    if rank != 0:
        return
    from oracle import loader as orc
    from pl_slam_plucker_b200 import abi, scene
    orc.build()
    cores = orc.set_threads(os.cpu_count() or 1)
    prof = PROFILES[args.profile]
    P = scene.make_scene(WORKLOADS[args.workload], line_mode=1 if prof == abi.PROFILE_H_END else 0)
    opt = abi.Options(prof, args.quirks)
    for _ in range(min(args.warmup, 1)):
        orc.solve(P, opt)
    t0 = time.perf_counter(); trials = 0
    steps = max(1, min(args.steps, 20))
    for _ in range(steps):
        r = orc.solve(P, opt); trials += r.n_trials
    dt = time.perf_counter() - t0
    val = P.n_obs * trials / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "lm_iters_per_s": trials / dt,
            "config": {"workload": workload_name(args, P), "profile": args.profile, "quirks": args.quirks, "windows_per_gpu": 1},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d full LBA solves of the %s window (oracle/plba_oracle.cpp, OpenMP over landmarks)" % (steps, args.workload)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


def _traffic(workload):
    """Measured DRAM bytes per k_assemble launch from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload]
        return int(t.get("k_assemble_w", t["k_assemble"]) if workload == "C5" else t["k_assemble"])
    except Exception:
        return None


def largest_config_roofline(s, abi, scene, args):
    """Assembly kernel on BASELINE config 5 (2 000 KFs, 2M points, 500k lines; inputs 0.45 GB >> L2, so every launch is cold)."""
    P5 = scene.make_scene(5)
    s.upload(P5, abi.Options(PROFILES[args.profile] if args.profile != "H_END" else 0, args.quirks))
    st = s.layout_stats(); kp = s.kernel_path()
    A = algorithmic_bytes(P5, st["nnzb_S"]); F = algorithmic_flops(P5)
    ms_a = s.time_kernel(0, 5); ms_u = s.time_kernel(2, 5); ms_s = s.time_kernel(1, 3)
    peak, src = _peak()
    return {"workload": "C5: %d free KFs, %d points, %d lines, %d+%d observations" % (P5.n_free, P5.n_pt, P5.n_ls, P5.n_pobs, P5.n_lobs),
            "kernel": "k_assemble_w" if kp["assembly"] == "warp" else "k_assemble", "kernel_path": kp, "bound": "hbm", "algorithmic_bytes": A, "ms_per_launch": ms_a, "achieved": A / (ms_a * 1e-3) / 1e9, "peak": peak,
            "peak_source": src, "unit": "GB/s", "frac": A / (ms_a * 1e-3) / 1e9 / peak, "traffic": _traffic("C5"), "algorithmic_flops": F, "achieved_fp64_tflops": F / (ms_a * 1e-3) / 1e12,
            "update_kernel_ms_per_launch": ms_u, "solve_ms_per_trial": ms_s, "lm_trial_ms": ms_a + ms_u + ms_s,
            "observations_per_s_per_trial": P5.n_obs / ((ms_a + ms_u + ms_s) * 1e-3), "nnzb_S": st["nnzb_S"], "launches_timed": 5,
            "note": "FP64 work (about %.1f GFLOP) bounds this kernel before HBM does: at the 37 TFLOP/s vector peak it needs %.0f us, the HBM roofline %.0f us" % (F / 1e9, F / 37e12 * 1e6, A / (peak * 1e9) * 1e6)}


def sharded_run(s, abi, scene, args, rank, world, dev, stream, barrier):
    """BASELINE config 4 sharded by base keyframe over the ranks, reduced camera system all-reduced with NCCL every LM trial."""
    import torch
    import torch.distributed as dist
    from pl_slam_plucker_b200 import sharded, solver
    P4 = scene.make_scene(WORKLOADS[args.sharded_workload])
    s2 = solver.LBASolver(dev.index, stream=stream.cuda_stream)
    sh = sharded.ShardedLBA(s2, rank, world, device=dev, nccl=True)       # the collective is the library's own ncclAllReduce (plba_comm_init_rank)
    opt = abi.Options(abi.PROFILE_G, 1)
    with torch.cuda.stream(stream):
        sh.upload(P4, opt)
        s2.run()                                   # warm-up (NCCL communicator, allocator)
        reps, t_ms, trials = 3, 0.0, 0
        for _ in range(reps):
            s2.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier(); e0.record(stream); s2.run(); e1.record(stream); torch.cuda.synchronize()
            t_ms += e0.elapsed_time(e1); trials += s2.timing()["n_trials_run"]
    t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_dbl = s2.reduced_system_ptr()[1]; kpath = s2.kernel_path()
    s2.close()
    return {"workload": "%s: %d free KFs, %d points, %d lines sharded by base keyframe over %d GPUs" % (args.sharded_workload, P4.n_free, P4.n_pt, P4.n_ls, world),
            "lm_trials": trials // reps, "ms_per_lba": float(t.item()) / reps, "observations_per_s": P4.n_obs * trials / (float(t.item()) * 1e-3),
            "allreduce_bytes_per_trial": int(8 * n_dbl), "kernel_path": kpath,
            "collective": "ncclAllReduce issued by the library on its own stream (plba_comm_init_rank), one per LM trial (+ 4 doubles after the update kernel), no host synchronisation inside the LM loop: the band of the reduced camera system in node form [D | U | b] when the block-cyclic-reduction solver runs, of the dense [S | g] otherwise"}


def workload_name(args, P):
    return "%s: %d free + %d fixed KFs, %d points, %d lines, %d+%d observations" % (
        args.workload, P.n_free, P.n_kf - P.n_free, P.n_pt, P.n_ls, P.n_pobs, P.n_lobs)


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
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the landmark-sharded run")
    ap.add_argument("--sharded-workload", default="C4", choices=["C4", "C5"], help="N > 1: the window that is sharded by base keyframe over the ranks")
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
    cfg = WORKLOADS[args.workload]
    sp = scene.preset(cfg)
    # weak scaling: every rank solves its own copy of the SAME synthetic window, so that the per-GPU work is identical (windows drawn with
    # different seeds take 19-25 LM trials and the step time, a maximum over ranks, would measure that imbalance instead)
    P = scene.make_scene(cfg, line_mode=1 if prof == abi.PROFILE_H_END else 0, seed=int(sp.seed))
    opt = abi.Options(prof, args.quirks)
    stream = torch.cuda.Stream(device=dev)
    s = solver.LBASolver(local, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-problem throughput -------------------------------------------------------------------------
    s.upload(P, opt)
    nnzb = s.layout_stats()["nnzb_S"]; kpath = s.kernel_path()
    for _ in range(args.warmup):
        s.reset(); s.run()
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    trials = launches = 0
    barrier()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                                              # L2 flush between timed iterations (untimed)
            s.reset()
            ev[k][0].record(stream)
            s.run()                                                    # the whole LM schedule: one CUDA-graph launch
            ev[k][1].record(stream)
        t = s.timing()
        launches += t["n_launches_run"]; trials += t["n_trials_run"]
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    # ---- per-kernel durations: the same steps replayed by the host-driven loop with CUDA events around each stage ----
    s.set_detail_timing(True)
    t_asm = t_sol = t_upd = 0.0; n_asm = 0
    det_steps = max(3, min(args.steps, 20))
    for k in range(det_steps):
        with torch.cuda.stream(stream):
            flush.zero_(); s.reset(); s.run()
        t = s.timing()
        t_asm += t["ms_assemble"]; t_sol += t["ms_solve"]; t_upd += t["ms_update"]; n_asm += t["n_assemble_run"]
    s.set_detail_timing(False)
    # ---- end to end through the drop-in call: host buffers in, host buffers out ----------------------------------
    out = abi.Result(P, 256)                       # caller-owned result buffers, re-used every call like a SLAM back-end would
    for _ in range(3):
        s.solve(P, opt, out=out)
    e2e_steps = max(3, args.steps // 4)
    barrier()
    t0 = time.perf_counter(); e2e_trials = 0; h2d = d2h = 0; hp = hu = gpu_ms = 0.0
    for _ in range(e2e_steps):
        r = s.solve(P, opt, out=out); e2e_trials += r.n_trials
        t = s.timing(); h2d += t["h2d_bytes"]; d2h += t["d2h_bytes"]; hp += t["ms_host_prep"]; hu += t["ms_host_unpack"]; gpu_ms += t["ms_total"]
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join()

    extras = {}
    if rank == 0 and not args.no_largest:
        extras["roofline_largest"] = largest_config_roofline(s, abi, scene, args)
    if world > 1 and not args.no_sharded:
        extras["sharded"] = sharded_run(s, abi, scene, args, rank, world, dev, stream, barrier)
    tt = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(P.n_obs * trials), float(P.n_obs * e2e_trials), float(trials)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = tt.tolist(); obs_trials, e2e_obs_trials, trials_all = cnt.tolist()

    if rank == 0:
        peak, peak_src = _peak()
        A = algorithmic_bytes(P, nnzb)
        asm_ms = t_asm / max(n_asm, 1)
        line = {"metric": METRIC, "value": obs_trials / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "lm_iters_per_s": trials_all / (ms_max * 1e-3), "lm_trials_per_step": trials / args.steps,
                "config": {"workload": workload_name(args, P), "profile": args.profile, "quirks": args.quirks, "windows_per_gpu": 1,
                           "l2": "flushed between steps (256 MiB write)", "parallelism": "one independent window per GPU (the same synthetic window on every rank), no collective"},
                "e2e": {"value": e2e_obs_trials / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                        "ms_per_step": e2e_ms_max / e2e_steps, "steps": e2e_steps,
                        "breakdown_ms": {"host_flatten": hp / e2e_steps, "device_lm_loop": gpu_ms / e2e_steps, "host_unpack": hu / e2e_steps}},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "kernel": ("k_assemble_w" if kpath["assembly"] == "warp" else "k_assemble") + " (one launch per LM trial: points + lines)", "kernel_path": kpath, "achieved": A / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else None,
                             "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": (A / (asm_ms * 1e-3) / 1e9 / peak) if asm_ms > 0 else None,
                             "traffic": _traffic(args.workload) if args.profile == "G" else None, "algorithmic_bytes": A, "ms_per_launch": asm_ms, "launches_timed": int(n_asm),
                             "stage_ms_per_step": {"assemble": t_asm / det_steps, "solve": t_sol / det_steps, "update": t_upd / det_steps},
                             "how": "CUDA events around each stage in a host-driven replay of the same steps (the headline steps run as one CUDA graph)",
                             "note": "BASELINE's metric quotes the roofline on the assembly stage, so this object is the assembly kernel; on this small window "
                                     "the largest share of the step is the reduced-system Cholesky (k_solve_small, %.0f %% of the stage time: n = %d sequential pivots, "
                                     "latency-bound, ~0.6 MFLOP), and every kernel is launch/latency-bound: see roofline_largest for the configuration whose "
                                     "assembly is throughput-bound" % (100.0 * t_sol / max(t_asm + t_sol + t_upd, 1e-9), 6 * P.n_free)}}
        line.update(extras)
        if not args.no_cpu_baseline:
            from oracle import loader as orc
            orc.build(); orc.set_threads(1)
            t0 = time.perf_counter(); n = 0; ctr = 0
            while time.perf_counter() - t0 < 10.0 or n < 1:
                ro = orc.solve(P, opt); ctr += ro.n_trials; n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": P.n_obs * ctr / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "%d full LBA solves of the same %s window by the oracle (single thread, as the reference: CMakeLists.txt:34 has no OpenMP), %.1f s" % (n, args.workload, dt),
                                    "host_cores_available": os.cpu_count()}
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
