"""Sustained A/B timing of k_run variants on the bench workload (16.7 M points x 365 days, M1 output into a 32-day ring).
   Variants (SEMIC_SHAPE values, read by the library at every launch: 24x0 = k_run_lean, 24x1 = the general k_run)
   are run in alternating blocks of `steps`
   launches so that power-cap clock drift hits all of them alike.
   python tools/ab.py 24x0,24x1 [k] [rounds] [steps] [points]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from semic_b200 import engine, surface_physics as sp
from semic_b200._capi import FIELD_ID, check, lib

shapes = sys.argv[1].split(",")
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
N = int(sys.argv[5]) if len(sys.argv) > 5 else 1 << 24
days = 365
L = lib()
st = sp.Surface_State_Class(); sp.surface_alloc(st, N)
check(L.semic_state_synth_mask(st.handle, 7, 0, 0.97, 0.02))
for name, v in (("tsurf", 260.0), ("hsnow", 1.0), ("alb", 0.8), ("alb_snow", 0.8)):
    check(L.semic_state_fill(st.handle, FIELD_ID[name], v))
forcing = engine.Series(N, 9, 73).synth_forcing(bench.SEED, daystride=5)
out = engine.Series(N, 8, 32)
bnd = sp.Boundary_Opt_Class()
par = sp.Surface_Param_Class()
for key, v in bench.par_dict("slater", k).items():
    setattr(par, key, v)
par.tsticsub = par.tstic / k
res = {s: [] for s in shapes}
for r in range(rounds):
    for s in shapes:
        os.environ["SEMIC_SHAPE"] = s
        ts = [engine.run(st, par, bnd, forcing, days, out=out, out_days=days) for _ in range(steps)]
        res[s] += ts[2:]
        print("round %d %s: %s" % (r, s, " ".join("%.1f" % t for t in ts)), flush=True)
for s in shapes:
    t = statistics.mean(res[s])
    print("k=%d %s: mean %.2f ms  %.2f G pt-days/s  frac %.3f" % (k, s, t, N * days / t / 1e6, N * days / t / 1e6 * 136 / 6458.7))
