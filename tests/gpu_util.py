"""Helpers for the -m gpu tests: build product-side objects from the same numpy inputs the oracle gets."""
import numpy as np

from oracle import binding as ob
from semic_b200 import engine
from semic_b200 import surface_physics as sp


def make_par(d):
    par = sp.Surface_Param_Class()
    for k, v in d.items():
        if k == "alb_scheme":
            par.alb_scheme = v
        elif k == "n_ksub":
            par.n_ksub = v
        else:
            setattr(par, k, v)
    par.tsticsub = par.tstic / float(par.n_ksub) if par.n_ksub else float("inf")    # surface_physics.f90:810
    return par


def make_bnd(names=()):
    bnd = sp.Boundary_Opt_Class()
    sp.surface_boundary_define(bnd, list(names))
    return bnd


def make_state(slab, mask):
    """Surface_State_Class holding a copy of an oracle slab (31, nx) + mask, uploaded to the device."""
    nx = slab.shape[1]
    st = sp.Surface_State_Class()
    sp.surface_alloc(st, nx)
    for n in ob.FIELDS:
        setattr(st, n, slab[ob.FIDX[n]])
    st.mask = mask
    st.upload()
    return st


def run_gpu(slab, mask, par_d, bnd_names, forcing, ndays, vali=None, out_days=0, strict=False, branch=True):
    """Runs semic_run on copies of the inputs.  Returns (final slab (31,nx), out [out_days][8][nx], branch uint8, ms)."""
    nx = slab.shape[1]
    par = make_par(par_d)
    bnd = make_bnd(bnd_names)
    st = make_state(slab, mask)
    fs = engine.Series(nx, 9, forcing.shape[0]).upload(forcing)
    vs = engine.Series(nx, 8, vali.shape[0]).upload(vali) if vali is not None else None
    outs = engine.Series(nx, 9 if branch else 8, out_days) if out_days > 0 else None
    ms = engine.run(st, par, bnd, fs, ndays, vali=vs, out=outs, out_days=out_days, strict=strict)
    st.download()
    final = st.as_slab()
    out = br = None
    if outs is not None:
        o = outs.download()
        out = np.ascontiguousarray(o[:, :8, :])
        if branch:
            br = o[:, 8, :].astype(np.uint8)
        outs.free()
    fs.free()
    if vs is not None:
        vs.free()
    sp.surface_dealloc(st)
    return final, out, br, ms


def run_oracle(slab, mask, par_d, bnd_names, forcing, ndays, vali=None, out_days=0):
    return run_oracle_diag(slab, mask, par_d, bnd_names, forcing, ndays, vali=vali, out_days=out_days)[:3]


def run_oracle_diag(slab, mask, par_d, bnd_names, forcing, ndays, vali=None, out_days=0):
    """-> (final slab, out, branch bitmap, r) with r = tmean/amp of the diurnal cycle per point-day (util.ill_conditioned)"""
    S = slab.copy()
    m = mask.copy()
    r = np.zeros((ndays, slab.shape[1]))
    out, br = ob.run(S, m, ob.make_par(par_d), ob.make_bnd(bnd_names), forcing, ndays, vali=vali,
                     out_days=out_days, want_branch=True, rdiag=r)
    if br is not None and out_days > 0:
        br = br[ndays - out_days:]
        r = r[ndays - out_days:]
    return S, out, br, r
