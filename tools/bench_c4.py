"""Config 4 (BASELINE.json configs[3]): isba albedo with surface_boundary_define overrides (tsurf / alb fed from an
external series, example.f90:103-107), 4 194 304 points, n_ksub=3; sustained timing of 365-day launches.
   python tools/bench_c4.py [points] [steps]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from semic_b200 import engine, surface_physics as sp
from semic_b200._capi import FIELD_ID, check, lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
days, W = 365, 73
L = lib()
st = sp.Surface_State_Class(); sp.surface_alloc(st, N)
check(L.semic_state_synth_mask(st.handle, 7, 0, 0.97, 0.02))
for name, v in (("tsurf", 260.0), ("hsnow", 1.0), ("alb", 0.8), ("alb_snow", 0.8)):
    check(L.semic_state_fill(st.handle, FIELD_ID[name], v))
forcing = engine.Series(N, 9, W).synth_forcing(bench.SEED, daystride=5)
out = engine.Series(N, 8, 32)
par = sp.Surface_Param_Class()
for key, v in bench.par_dict("isba", 3).items():
    setattr(par, key, v)
par.tsticsub = par.tstic / 3
# external fields: the model's own daily output of a plain run over the window (plausible tsurf / alb series)
vali = engine.Series(N, 8, W)
bnd0 = sp.Boundary_Opt_Class()
engine.run(st, par, bnd0, forcing, W, out=vali, out_days=W)
sets = [(), ("tsurf",), ("tsurf", "alb")] + ([("melt",), ("acc", "smb"), ("shf", "lhf"), ("tsurf", "alb", "melt", "acc", "smb")] if os.environ.get("C4_ALL") else [])
for names in sets:
    bnd = sp.Boundary_Opt_Class()
    sp.surface_boundary_define(bnd, list(names))
    ts = [engine.run(st, par, bnd, forcing, days, vali=vali if names else None, out=out, out_days=days) for _ in range(steps)]
    t = statistics.mean(ts[2:])
    nbytes = 136 + 8 * len([n for n in names if n in ('tsurf', 'alb', 'melt', 'acc', 'smb')])
    rate = N * days / t / 1e6
    print("C4 isba overrides=%-14s %.2f ms  %.2f G pt-days/s  %d B/pt-day -> %.0f GB/s (%.3f of the HBM roof)"
          % ("+".join(names) or "none", t, rate, nbytes, rate * nbytes, rate * nbytes / 6458.7), flush=True)
