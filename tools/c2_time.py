"""BASELINE configs[1] timing: 100 000 points x 365 days, n_ksub = 24, slater, forcing year resident, all daily rows written.
   python tools/c2_time.py [points] [n_ksub]        (SEMIC_B200_LIB selects the library build)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from semic_b200 import engine, surface_physics as sp
from semic_b200._capi import FIELD_ID, check, lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
L = lib()
st = sp.Surface_State_Class(); sp.surface_alloc(st, N)
check(L.semic_state_synth_mask(st.handle, 7, 0, 0.85, 0.10))
for name, v in (("tsurf", 260.0), ("hsnow", 1.0), ("alb", 0.8), ("alb_snow", 0.8)):
    check(L.semic_state_fill(st.handle, FIELD_ID[name], v))
forcing = engine.Series(N, 9, 365).synth_forcing(bench.SEED)
out = engine.Series(N, 8, 365)
bnd = sp.Boundary_Opt_Class()
par = sp.Surface_Param_Class()
for key, v in bench.par_dict("slater", K).items():
    setattr(par, key, v)
par.tsticsub = par.tstic / K
ts = [engine.run(st, par, bnd, forcing, 365, out=out, out_days=365) for _ in range(12)][2:]
t = statistics.mean(ts)
print("n_ksub %d, %d points: mean %.4f ms per year, min %.4f  ->  %.2f G point-days/s" % (K, N, t, min(ts), N * 365 / t / 1e6))
