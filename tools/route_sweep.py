"""Where the size routing between the CTA-chunk and the warp kernels should sit: whole-LBA time of batches of C2-shaped windows."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_slam_plucker_b200 import abi, scene, solver
s = solver.LBASolver(0)
opt = abi.Options(abi.PROFILE_G, 0)
for nwin in (1, 2, 3, 4, 6, 8, 16):
    probs = scene.make_batch(nwin, 2)
    row = {"windows": nwin, "obs": sum(p.n_obs for p in probs)}
    for path in (1, 2):
        s.set_kernel_path(path); s.upload(probs, opt)
        ts = []
        for _ in range(5):
            s.reset(); t = time.perf_counter(); s.run(); ts.append(time.perf_counter() - t)
        row["chunk_ms" if path == 1 else "warp_ms"] = round(1e3 * min(ts), 3)
    print(json.dumps(row), flush=True)
