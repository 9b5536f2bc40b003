import sys, os, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from pl_slam_plucker_b200 import abi, scene, solver, _lib
from oracle import loader as orc
from helpers import assert_trace_close, assert_state_close
orc.set_threads(os.cpu_count())
probs = scene.make_batch(1024, 3); opt = abi.Options(abi.PROFILE_G, 0)
lib = _lib.load(sys.argv[1]) if len(sys.argv) > 1 else None
s = solver.LBASolver(0, lib=lib)
rc, rs = s.solve_batch(probs, opt)
orc_rc, os_ = orc.solve_batch(probs, opt)
bad=0
for w,(P,r,o) in enumerate(zip(probs,rs,os_)):
    try:
        assert_trace_close(o.trace, r.trace, abi.PROFILE_G); assert_state_close(o, r, P, abi.PROFILE_G)
    except AssertionError as e:
        bad+=1
        if bad<6:
            n=min(len(o.trace),len(r.trace))
            print("window",w,"len",len(o.trace),len(r.trace),str(e)[:300].replace("\n"," | "))
            k=[i for i in range(n) if o.trace['accepted'][i]!=r.trace['accepted'][i] or o.trace['iter'][i]!=r.trace['iter'][i]]
            print("  first decision diff",k[:1], "rho", o.trace['rho'][k[0]] if k else None, r.trace['rho'][k[0]] if k else None, "pose", np.abs(o.kf_T_wc-r.kf_T_wc).max())
print("bad windows",bad, "lib", sys.argv[1:] )
mx=[max(float(np.max(np.abs(r.trace["chi"][:min(len(r.trace),len(o.trace))]-o.trace["chi"][:min(len(r.trace),len(o.trace))])/np.abs(o.trace["chi"][:min(len(r.trace),len(o.trace))]))),0) for r,o in zip(rs,os_)]
print("max rel chi dev over windows: max %.2e, 99th pct %.2e, median %.2e, count>1e-9: %d" % (max(mx), np.percentile(mx,99), np.median(mx), sum(m>1e-9 for m in mx)))
