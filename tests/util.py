"""Shared helpers of the test-suite: parameter sets, initial states, comparisons, flip report."""
import os

import numpy as np

from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# example/example.namelist:10-31 (values restated in tests/golden/example.nml)
PAR_EXAMPLE = dict(tstic=86400., ceff=2.0e6, csh=2.0e-3, clh=5.0e-4, alb_smax=0.81, alb_smin=-999., albi=0.07,
                   albl=0.15, tmin=-999., tmax=273.15, hcrit=0.028, rcrit=0.79, amp=3.5, alb_scheme="none",
                   tau_a=0.008, tau_f=0.24, w_crit=15.0, mcrit=6.0e-8, afac=0.0, tmid=0.0, n_ksub=3)
# f2py/test.nml with rcrit=0.5 (the shipped -999 makes f_rz = -999)
PAR_SLATER = dict(tstic=86400., ceff=2.0e6, csh=2.0e-3, clh=5.0e-4, alb_smax=0.80, alb_smin=0.77, albi=0.45,
                  albl=0.15, tmin=263.15, tmax=273.15, hcrit=0.09, rcrit=0.5, amp=3.1, alb_scheme="slater",
                  tau_a=0.008, tau_f=0.24, w_crit=15.0, mcrit=6.0e-8, afac=-0.18, tmid=275.35, n_ksub=3)
PAR_ISBA = dict(PAR_SLATER, alb_scheme="isba", alb_smin=0.6, alb_smax=0.85)
PAR_DENBY = dict(PAR_SLATER, alb_scheme="denby", alb_smin=0.6, alb_smax=0.85)
PAR_ALEX = dict(PAR_SLATER, alb_scheme="alex", alb_smin=0.6, alb_smax=0.85)     # afac/tmid as optimize/tools.py:28-29
PAR_UNKNOWN = dict(PAR_SLATER, alb_scheme="foo")
SCHEME_PARS = {"none": dict(PAR_SLATER, alb_scheme="none"), "slater": PAR_SLATER, "isba": PAR_ISBA,
               "denby": PAR_DENBY, "alex": PAR_ALEX, "unknown": PAR_UNKNOWN}

# absolute floors of the relative error per variable: |gpu-ref| / max(|ref|, floor).  A floor is ~1e-3 of the
# variable's typical magnitude in the shipped forcing, so that exact zeros and denormal-size leftovers of
# cancellations do not demand 1e-10 of nothing.
FLOOR = dict(t2m=1.0, tsurf=1.0, hsnow=1e-3, hice=1e-3, alb=1e-2, alb_snow=1e-2, melt=1e-10, melted_snow=1e-10,
             melted_ice=1e-10, refr=1e-10, smb=1e-10, acc=1e-10, lhf=1e-2, shf=1e-2, lwu=1.0, subl=1e-10, evap=1e-8,
             smb_snow=1e-10, smb_ice=1e-10, runoff=1e-10, qmr=1e-1, qmr_res=1e-1, amp=1.0,
             sf=1e-10, rf=1e-10, sp=1.0, lwd=1.0, swd=1.0, wind=1.0, rhoa=1.0, qq=1e-6)
OUT_FLOOR = [FLOOR["tsurf"], FLOOR["alb"], 1e-1, FLOOR["smb"], FLOOR["melt"], FLOOR["acc"], FLOOR["shf"], FLOOR["lhf"]]
OUT_NAMES = ["stt", "alb", "swnet", "smb", "melt", "acc", "shf", "lhf"]
TOL = 1e-10          # north_star: FP64 state within 1e-10 relative per variable after one simulated year


def load_fixture(name):
    forc = np.load(os.path.join(GOLDEN, name + "_input.npy"))
    vali = np.load(os.path.join(GOLDEN, name + "_vali.npy"))
    return forc, vali


def random_forcing(nx, ndays, seed=1):
    """numpy synthetic forcing in the ranges of the shipped data (CPU-side tests only)."""
    r = np.random.default_rng(seed)
    tbar = r.uniform(245., 272., nx)
    doy = np.arange(ndays)[:, None]
    t2m = tbar[None, :] + 12. * np.cos(2 * np.pi * (doy - 200) / 365.) + 3. * r.standard_normal((ndays, nx))
    swd = np.maximum(0., 400. * np.sin(np.pi * ((doy % 365) - 35) / 295.)) * r.uniform(0.6, 1., (ndays, nx))
    lwd = np.clip(200. + 4. * (t2m - 260.) + 15. * r.standard_normal((ndays, nx)), 70., 340.)
    wind = np.clip(6.5 + 3. * r.standard_normal((ndays, nx)), 0.1, 22.)
    sp = r.uniform(63000., 100000., nx)[None, :] + 1200. * r.standard_normal((ndays, nx))
    rhoa = sp / (287.05 * t2m)
    x = t2m - 273.15
    es = np.where(t2m < 273.15, 611.2 * np.exp(22.46 * x / (272.62 + x)), 611.2 * np.exp(17.62 * x / (243.12 + x)))
    qq = r.uniform(0.5, 0.95, (ndays, nx)) * es * 0.62197 / (es * (0.62197 - 1.) + sp)
    sf = np.where((t2m < 273.15) & (r.uniform(size=(ndays, nx)) > 0.65), r.exponential(3.5e-8, (ndays, nx)), 0.)
    rf = np.where((t2m > 268.) & (r.uniform(size=(ndays, nx)) > 0.6), r.exponential(6e-9, (ndays, nx)), 0.)
    f = np.stack([sf, rf, swd, lwd, wind, sp, rhoa, qq, t2m], axis=1)       # [day][9][nx]
    return np.ascontiguousarray(f)


def init_slab(nx, mask, tsurf=260., hsnow=1.0, alb=0.8, hice=0.0):
    S, m = ob.new_state(nx)
    m[:] = mask
    S[ob.FIDX["tsurf"]] = tsurf
    S[ob.FIDX["hsnow"]] = hsnow
    S[ob.FIDX["hice"]] = hice
    S[ob.FIDX["alb"]] = alb
    S[ob.FIDX["alb_snow"]] = alb
    return S, m


def rel_err(a, b, floor):
    """pointwise |a-b| / max(|b|, floor) (supplementary, ill-conditioned near cancellations)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bad = ~(np.isfinite(a) & np.isfinite(b))
    same_nonfinite = bad & ((np.isnan(a) & np.isnan(b)) | (a == b))
    d = np.abs(a - b) / np.maximum(np.abs(b), floor)
    d = np.where(same_nonfinite, 0.0, d)
    d = np.where(bad & ~same_nonfinite, np.inf, d)
    return d


def norm_err(a, b, floor, exclude=None):
    """THE parity metric: relative error of a variable as a field, max|a-b| / max(max|b|, floor).

    `exclude` masks entries that are reported separately (point-days after a threshold flip); the
    denominator is the variable's magnitude over the whole field."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    bad = ~(np.isfinite(a) & np.isfinite(b))
    same_nonfinite = bad & ((np.isnan(a) & np.isnan(b)) | (a == b))
    d = np.where(same_nonfinite, 0.0, np.abs(np.where(bad, 0.0, a) - np.where(bad, 0.0, b)))
    d = np.where(bad & ~same_nonfinite, np.inf, d)
    if exclude is not None:
        d = np.where(exclude, 0.0, d)
    fin = np.abs(b[np.isfinite(b)])
    scale = max(float(fin.max()) if fin.size else 0.0, floor)
    return float(d.max()) / scale


def state_errors(slab_gpu, slab_ref, names=None, exclude_points=None, pointwise=False):
    """relative error per state variable (fields 0..22); returns dict name -> err"""
    out = {}
    for n in (names or ob.FIELDS[:23]):
        if pointwise:
            e = rel_err(slab_gpu[ob.FIDX[n]], slab_ref[ob.FIDX[n]], FLOOR[n])
            if exclude_points is not None and e.size:
                e = np.where(exclude_points, 0.0, e)
            out[n] = float(e.max()) if e.size else 0.0
        else:
            out[n] = norm_err(slab_gpu[ob.FIDX[n]], slab_ref[ob.FIDX[n]], FLOOR[n], exclude_points)
    return out


def out_errors(out_gpu, out_ref, exclude=None, pointwise=False):
    """out arrays [days][8][nx]; exclude: bool [days][nx] of point-days to leave out. dict name -> err"""
    res = {}
    for v, n in enumerate(OUT_NAMES):
        if pointwise:
            e = rel_err(out_gpu[:, v, :], out_ref[:, v, :], OUT_FLOOR[v])
            if exclude is not None and e.size:
                e = np.where(exclude, 0.0, e)
            res[n] = float(e.max()) if e.size else 0.0
        else:
            res[n] = norm_err(out_gpu[:, v, :], out_ref[:, v, :], OUT_FLOOR[v], exclude)
    return res


HSNOW_ZERO = 64     # the "hsnow == 0" bit: a one-ulp ghost snow layer is immaterial unless it flips another decision


def flip_report(branch_gpu, branch_ref):
    """branch arrays [days][nx] uint8.  Returns (n_flipped_pointdays, flipped_points mask [nx], first-flip day [nx]).
    Counts MATERIAL flips (any decision other than the hsnow==0 indicator)."""
    diff = ((branch_gpu.astype(np.uint8) ^ branch_ref.astype(np.uint8)) & ~np.uint8(HSNOW_ZERO)) != 0
    pts = diff.any(axis=0)
    first = np.where(pts, diff.argmax(axis=0), -1)
    return int(diff.sum()), pts, first


def flips_by_bit(branch_gpu, branch_ref):
    d = branch_gpu.astype(np.uint8) ^ branch_ref.astype(np.uint8)
    names = ["ts<t0", "clamp", "gate", "diurnal_lo", "diurnal_hi", "melt>0", "hsnow==0", "snow2ice"]
    return {n: int(((d >> b) & 1).sum()) for b, n in enumerate(names)}


def tainted(branch_gpu, branch_ref):
    """bool [days][nx]: point-days at or after the first MATERIAL threshold flip of that point (reported separately)."""
    diff = ((branch_gpu.astype(np.uint8) ^ branch_ref.astype(np.uint8)) & ~np.uint8(HSNOW_ZERO)) != 0
    return np.maximum.accumulate(diff, axis=0)
