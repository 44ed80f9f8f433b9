"""ctypes binding of the CPU oracle (oracle/semic_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under semic_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsemic_oracle.so")

NFIELDS = 31
FIELDS = ["t2m", "tsurf", "hsnow", "hice", "alb", "alb_snow", "melt", "melted_snow", "melted_ice",
          "refr", "smb", "acc", "lhf", "shf", "lwu", "subl", "evap", "smb_snow", "smb_ice", "runoff",
          "qmr", "qmr_res", "amp", "sf", "rf", "sp", "lwd", "swd", "wind", "rhoa", "qq"]
FIDX = {n: i for i, n in enumerate(FIELDS)}
PAR_DOUBLES = ["ceff", "albi", "albl", "alb_smax", "alb_smin", "hcrit", "rcrit", "amp", "csh", "clh",
               "tmin", "tmax", "tstic", "tsticsub", "tau_a", "tau_f", "w_crit", "mcrit", "afac", "tmid"]
BND_NAMES = ["t2m", "tsurf", "hsnow", "alb", "melt", "refr", "smb", "acc", "lhf", "shf", "subl", "amp"]


class OraclePar(C.Structure):
    _fields_ = [("alb_scheme", C.c_char * 256), ("nx", C.c_int), ("n_ksub", C.c_int)] + \
               [(n, C.c_double) for n in PAR_DOUBLES]


class OracleBnd(C.Structure):
    _fields_ = [(n, C.c_int) for n in BND_NAMES]


class OracleState(C.Structure):
    """oracle_state of semic_oracle.h (for tests that call oracle_energy_balance / oracle_mass_balance directly)"""
    _fields_ = [("nx", C.c_int), ("ld", C.c_int), ("slab", C.POINTER(C.c_double)), ("mask", C.POINTER(C.c_int)),
                ("branch", C.POINTER(C.c_ubyte)), ("rdiag", C.POINTER(C.c_double))]


def build(force=False):
    src = os.path.join(_HERE, "semic_oracle.c")
    hdr = os.path.join(_HERE, "semic_oracle.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libsemic_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        bp = C.POINTER(C.c_ubyte)
        L.oracle_step.argtypes = [dp, C.c_int, ip, C.c_int, C.POINTER(OraclePar), C.POINTER(OracleBnd), bp]
        L.oracle_step.restype = None
        L.oracle_run.argtypes = [dp, C.c_int, ip, C.c_int, C.POINTER(OraclePar), C.POINTER(OracleBnd),
                                 dp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, bp]
        L.oracle_run.restype = None
        L.oracle_run_diag.argtypes = L.oracle_run.argtypes + [dp]
        L.oracle_run_diag.restype = None
        for name, nargs in [("oracle_sensible_heat_flux", 6), ("oracle_longwave_upward", 2),
                            ("oracle_ew_sat", 1), ("oracle_ei_sat", 1), ("oracle_albedo_slater", 5),
                            ("oracle_albedo_denby", 4)]:
            f = getattr(L, name)
            f.argtypes = [C.c_double] * nargs
            f.restype = C.c_double
        L.oracle_latent_heat_flux.argtypes = [C.c_double] * 5 + [C.c_int] + [C.c_double] * 4 + [dp] * 3
        L.oracle_latent_heat_flux.restype = None
        L.oracle_diurnal_cycle.argtypes = [C.c_double, C.c_double, dp, dp]
        L.oracle_diurnal_cycle.restype = None
        L.oracle_albedo_isba.argtypes = [dp] + [C.c_double] * 10
        L.oracle_albedo_isba.restype = None
        L.oracle_field_average.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double]
        L.oracle_field_average.restype = C.c_int
        L.oracle_surface_physics_average.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_double]
        L.oracle_surface_physics_average.restype = C.c_int
        L.oracle_calculate_crmsd.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, dp]
        L.oracle_calculate_crmsd.restype = None
        L.oracle_particle_cost.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int]
        L.oracle_particle_cost.restype = C.c_double
        L.oracle_particle_cost_sumsq.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int]
        L.oracle_particle_cost_sumsq.restype = C.c_double
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def make_par(d):
    """dict with the namelist keys (surface_physics.f90:743-749) -> OraclePar; tsticsub per :810."""
    p = OraclePar()
    p.alb_scheme = str(d.get("alb_scheme", "")).encode()
    p.nx = int(d.get("nx", 0))
    p.n_ksub = int(d["n_ksub"])
    for n in PAR_DOUBLES:
        if n != "tsticsub":
            setattr(p, n, float(d.get(n, 0.0)))
    p.tsticsub = p.tstic / float(p.n_ksub)
    return p


def make_bnd(names=()):
    """surface_boundary_define (surface_physics.f90:819-878): unknown names are ignored."""
    b = OracleBnd()
    for n in names:
        n = n.strip()
        if n in BND_NAMES:
            setattr(b, n, 1)
    return b


def new_state(nx):
    """Zero-filled (31, nx) slab + int32 mask (Q8: the new alloc zero-fills)."""
    return np.zeros((NFIELDS, nx), dtype=np.float64), np.zeros(nx, dtype=np.int32)


def step(slab, mask, par, bnd, branch=None):
    nx = slab.shape[1]
    assert slab.flags.c_contiguous and slab.dtype == np.float64 and mask.dtype == np.int32
    lib().oracle_step(_dp(slab), slab.shape[1], mask.ctypes.data_as(C.POINTER(C.c_int)), nx,
                      C.byref(par), C.byref(bnd),
                      branch.ctypes.data_as(C.POINTER(C.c_ubyte)) if branch is not None else None)


def run(slab, mask, par, bnd, forcing, ndays, vali=None, out_days=0, nout=None, want_branch=False, nx=None, rdiag=None):
    """forcing [nwin][9][ldf], vali [nwin][8][ldf] or None.  Returns (out [nout][8][nx] or None, branch or None).
    rdiag: optional float64 [ndays][nx] that receives tmean/amp of every point-day (conditioning report)."""
    nx = slab.shape[1] if nx is None else nx
    nwin, nine, ldf = forcing.shape
    assert nine == 9 and forcing.flags.c_contiguous and forcing.dtype == np.float64
    if vali is not None:
        assert vali.shape == (nwin, 8, ldf) and vali.flags.c_contiguous
    out = None
    if out_days > 0:
        nout = out_days if nout is None else nout
        out = np.zeros((nout, 8, nx), dtype=np.float64)
    branch = np.zeros((ndays, nx), dtype=np.uint8) if want_branch else None
    if rdiag is not None:
        assert rdiag.shape == (ndays, nx) and rdiag.dtype == np.float64 and rdiag.flags.c_contiguous
    lib().oracle_run_diag(_dp(slab), slab.shape[1], mask.ctypes.data_as(C.POINTER(C.c_int)), nx,
                          C.byref(par), C.byref(bnd), _dp(forcing), _dp(vali), nwin, ldf, ndays,
                          _dp(out), (nout or 1), nx, out_days,
                          branch.ctypes.data_as(C.POINTER(C.c_ubyte)) if branch is not None else None, _dp(rdiag))
    return out, branch


def calculate_crmsd(x1, x2):
    """x1, x2: [ntime][nx]"""
    ntime, nx = x1.shape
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    r = np.zeros(nx)
    lib().oracle_calculate_crmsd(_dp(x1), _dp(x2), ntime, nx, nx, _dp(r))
    return r


def particle_cost(out, vali, sumsq=False):
    """out, vali: [ntime][8][nx]"""
    ntime, eight, nx = out.shape
    out = np.ascontiguousarray(out, dtype=np.float64)
    vali = np.ascontiguousarray(vali, dtype=np.float64)
    f = lib().oracle_particle_cost_sumsq if sumsq else lib().oracle_particle_cost
    return f(_dp(out), _dp(vali), ntime, nx, nx)
