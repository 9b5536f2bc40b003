"""Config 3 (batch of independent 10-KF windows) throughput: python tools/batch_bench.py [n_windows]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t = time.time(); probs = scene.make_batch(n, 3); tg = time.time() - t
s = solver.LBASolver(0)
opt = abi.Options(abi.PROFILE_G, 0)
t = time.time(); s.upload(probs, opt); tu = time.time() - t
nobs = sum(p.n_obs for p in probs)
ts = []
for _ in range(3):
    s.reset(); t = time.time(); s.run(); ts.append(time.time() - t)
tm = s.timing()
print(json.dumps({"windows": n, "obs": nobs, "gen_s": round(tg, 2), "upload_s": round(tu, 3), "run_ms": round(1e3 * min(ts), 2), "trials": tm["n_trials_run"],
                  "launches": tm["n_launches_run"], "obs_trials_per_s": nobs / n * tm["n_trials_run"] / min(ts), "windows_per_s": n / min(ts),
                  "assemble_ms": s.time_kernel(0, 3), "solve_ms": s.time_kernel(1, 3), "update_ms": s.time_kernel(2, 3)}))
