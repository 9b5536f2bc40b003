"""Pose-tracking throughput (plba_track_solve): a batch of independent synthetic frames on the GPU against the CPU oracle.
    python tools/track_bench.py [n_frames]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pl_slam_plucker_b200 import solver, tracking as trk
from oracle import loader as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cam = (435.2, 435.2, 367.2, 252.2)
frames = [trk.make_frame(i)[0] for i in range(n)]
opt = trk.Options(cam, max_iters=5)
s = solver.LBASolver(0)
trk.solve(s, frames[:8], opt)
out = {}
for m in (1, 64, n):
    b = trk.Batch(frames[:m]); ts = []
    for _ in range(5):
        t = time.perf_counter(); trk.solve(s, b, opt, unpack=False); ts.append(time.perf_counter() - t)     # the C call: host packing + H2D + kernel + D2H
    out["gpu_ms_batch_%d" % m] = round(1e3 * min(ts), 3)
res = trk.solve(s, frames, opt)
out["gpu_frames_per_s"] = n / (out["gpu_ms_batch_%d" % n] * 1e-3)
t = time.perf_counter(); k = 0
while time.perf_counter() - t < 3.0:
    orc.track_solve(frames[k % n], opt); k += 1
out["cpu_oracle_frames_per_s_1_thread"] = k / (time.perf_counter() - t)
out["matches_per_frame"] = "200 points + 80 lines"; out["iters"] = int(np.mean([r["iters"] for r in res]))
print(json.dumps(out))
