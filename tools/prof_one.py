"""One configuration of k_run for ncu: python tools/prof_one.py <points> <days> <k> [unused] [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from semic_b200 import engine, surface_physics as sp
from semic_b200._capi import FIELD_ID, check, lib

N = int(sys.argv[1]); days = int(sys.argv[2]); k = int(sys.argv[3])
nl = int(sys.argv[5]) if len(sys.argv) > 5 else 3
L = lib()
st = sp.Surface_State_Class(); sp.surface_alloc(st, N)
check(L.semic_state_synth_mask(st.handle, 7, 0, 0.97, 0.02))
for name, v in (("tsurf", 260.0), ("hsnow", 1.0), ("alb", 0.8), ("alb_snow", 0.8)):
    check(L.semic_state_fill(st.handle, FIELD_ID[name], v))
W = min(days, 73)
forcing = engine.Series(N, 9, W).synth_forcing(bench.SEED, daystride=5 if W == 73 else 1)
out = engine.Series(N, 8, 32)
bnd = sp.Boundary_Opt_Class()
par = sp.Surface_Param_Class()
for key, v in bench.par_dict("slater", k).items():
    setattr(par, key, v)
par.tsticsub = par.tstic / k
for i in range(nl):
    t = engine.run(st, par, bnd, forcing, days, out=out, out_days=days)
    print("launch %d: %.3f ms  %.2f G pt-days/s" % (i, t, N * days / (t * 1e-3) / 1e9), flush=True)
