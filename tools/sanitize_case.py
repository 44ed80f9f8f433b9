"""A small tour of every kernel and entry point (meant for compute-sanitizer, which is closed on this pool; plain runs only):
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from semic_b200 import driver, engine, surface_physics as sp, synth

nx, nd = 300, 12
forc = synth.forcing(nx, nd)
mask = synth.mask(nx)
for k, scheme, names in ((3, "slater", ()), (24, "slater", ()), (4, "isba", ("tsurf", "alb")), (3, "denby", ("melt", "acc"))):
    par = sp.Surface_Param_Class()
    for key, v in bench.par_dict(scheme, k).items():
        setattr(par, key, v)
    par.tsticsub = par.tstic / k
    bnd = sp.Boundary_Opt_Class()
    sp.surface_boundary_define(bnd, list(names))
    st = sp.Surface_State_Class(); sp.surface_alloc(st, nx)
    st.mask = mask; st.tsurf = 260.0; st.hsnow = 1.0; st.alb = 0.8; st.alb_snow = 0.8
    st.upload()
    fs = engine.Series(nx, 9, nd).upload(forc)
    out = engine.Series(nx, 8, 5)
    vs = engine.Series(nx, 8, nd)
    engine.run(st, par, sp.Boundary_Opt_Class(), fs, nd, out=vs, out_days=nd)             # something to override with
    engine.run(st, par, bnd, fs, 2 * nd, vali=vs if names else None, out=out, out_days=7)   # window wraps, ring wraps
    if not names:
        engine.run(st, par, bnd, fs, nd, out=out, out_days=nd, avg_days=4)
        engine.run(st, par, bnd, fs, nd, out=out, out_days=5, strict=True)
        o9 = engine.Series(nx, 9, nd)
        engine.run(st, par, bnd, fs, nd, out=o9, out_days=nd)                                # bitmap kernel
        sp.surface_energy_and_mass_balance(st, par, bnd, 1, 0)                               # drop-in single day
        hf = engine.PinnedArray((nd, 9, nx)); hf.array[...] = forc
        ho = engine.PinnedArray((nd, 8, nx))
        engine.run_host(st, par, bnd, hf.array, nd, h_out=ho.array, out_days=nd, chunk_days=5)
    print("ok", k, scheme, names, flush=True)
    sp.surface_dealloc(st)
# ensemble
names = ["hcrit", "alb_smax", "amp", "rcrit", "albi", "albl"]
r = np.random.default_rng(0)
pop = [{"id": i, "pos": list(np.array([0.05, 0.8, 1.0, 0.2, 0.4, 0.1]) + 0.05 * r.uniform(size=6))} for i in range(11)]
vali = np.ascontiguousarray(r.normal(size=(nd, 8, nx)) + 1.0)
cost = driver.run_particles(pop, names, forc, vali, mask, dict(hsnow=1.0, hice=0.0, alb=0.8, tsurf=260.0, alb_snow=0.8), nd, 2)
print("ok ensemble", len(cost))
