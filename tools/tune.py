"""Times k_run variants (SEMIC_SHAPE x n_ksub) on one allocation.  python tools/tune.py [points] [days]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from semic_b200 import engine, surface_physics as sp
from semic_b200._capi import FIELD_ID, check, lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
days = int(sys.argv[2]) if len(sys.argv) > 2 else 73
L = lib()
st = sp.Surface_State_Class(); sp.surface_alloc(st, N)
check(L.semic_state_synth_mask(st.handle, 7, 0, 0.97, 0.02))
for name, v in (("tsurf", 260.0), ("hsnow", 1.0), ("alb", 0.8), ("alb_snow", 0.8)):
    check(L.semic_state_fill(st.handle, FIELD_ID[name], v))
forcing = engine.Series(N, 9, 73).synth_forcing(bench.SEED, daystride=5)
out = engine.Series(N, 8, 32)
bnd = sp.Boundary_Opt_Class()
for k in (3, 24):
    par = sp.Surface_Param_Class()
    for key, v in bench.par_dict("slater", k).items():
        setattr(par, key, v)
    par.tsticsub = par.tstic / k
    for shape in (os.environ.get("TUNE_SHAPES", "24x0,24x1").split(",")):
        os.environ["SEMIC_SHAPE"] = shape
        engine.run(st, par, bnd, forcing, 8, out=out, out_days=8)
        ts = [engine.run(st, par, bnd, forcing, days, out=out, out_days=days) for _ in range(3)]
        t = min(ts)
        rate = N * days / (t * 1e-3)
        print("k=%2d shape=%s: %.2f ms  %.2f G pt-days/s  %.0f GB/s (M1 136 B)  no-out:" % (k, shape, t, rate / 1e9, rate * 136 / 1e9), end=" ")
        t0 = min(engine.run(st, par, bnd, forcing, days) for _ in range(2))
        print("%.2f G pt-days/s" % (N * days / (t0 * 1e-3) / 1e9), flush=True)
