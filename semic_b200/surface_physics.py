"""Python mirror of `module surface_physics` over the C-ABI (include/semic_b200.h).

Same names and argument meaning as the f90wrap-generated module the reference builds
(f2py/Makefile:127-134; usage f2py/test.py:14-75):

    import SurfacePhysics as sp
    par = sp.surface_physics.Surface_Param_Class()
    sp.surface_physics.surface_physics_par_load(par, 'test.nml')
    bnd = sp.surface_physics.Boundary_Opt_Class()
    sp.surface_physics.surface_boundary_define(bnd, par.boundary)
    state = sp.surface_physics.Surface_State_Class()
    sp.surface_physics.surface_alloc(state, par.nx)
    state.tsurf = ...                      # array attributes are numpy views; assignment copies in
    sp.surface_physics.surface_energy_and_mass_balance(state, par, bnd, day, year)

Everything that computes runs on the GPU through libsemic_b200.so; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (BND_NAMES, FIELD_ID, FIELD_NAMES, NBOUNDARY, PAR_DOUBLES, STRLEN, SemicBnd, SemicError,
                    SemicPar, check, fstr, lib, padded)

# public constants of the module (surface_physics.f90:129)
sigm = 5.67e-8
cap = 1000.0
eps = 0.62197
cls = 2.83e6
clv = 2.5e6


class Surface_Param_Class(object):
    """surface_param_class (surface_physics.f90:28-54)."""

    def __init__(self):
        object.__setattr__(self, "_c", SemicPar())
        # Fortran leaves the components undefined; start from blanks/zeros
        self._c.name = b""
        self._c.alb_scheme = b""

    def __getattr__(self, k):
        c = object.__getattribute__(self, "_c")
        if k in PAR_DOUBLES or k in ("nx", "n_ksub"):
            return getattr(c, k)
        if k in ("name", "alb_scheme"):
            return fstr(getattr(c, k))
        if k == "boundary":
            return [fstr(bytes(c.boundary[q])) for q in range(NBOUNDARY)]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        c = object.__getattribute__(self, "_c")
        if k in PAR_DOUBLES:
            setattr(c, k, float(v))
        elif k in ("nx", "n_ksub"):
            setattr(c, k, int(v))          # f2py/test.py:16 assigns a float under Python 2 semantics
        elif k in ("name", "alb_scheme"):
            C.memmove(C.addressof(c) + getattr(SemicPar, k).offset, padded(v), STRLEN)
        elif k == "boundary":
            vals = list(v) + [""] * (NBOUNDARY - len(v))
            for q in range(NBOUNDARY):
                C.memmove(C.addressof(c) + SemicPar.boundary.offset + q * STRLEN, padded(vals[q]), STRLEN)
        else:
            object.__setattr__(self, k, v)


class Boundary_Opt_Class(object):
    """boundary_opt_class (surface_physics.f90:92-105)."""

    def __init__(self):
        object.__setattr__(self, "_c", SemicBnd())

    def __getattr__(self, k):
        if k in BND_NAMES:
            return bool(getattr(object.__getattribute__(self, "_c"), k))
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if k in BND_NAMES:
            setattr(object.__getattribute__(self, "_c"), k, 1 if v else 0)
        else:
            object.__setattr__(self, k, v)


class Surface_State_Class(object):
    """surface_state_class (surface_physics.f90:56-90): 31 double arrays + integer mask.

    Array attributes are numpy views of the pinned host mirror the device state is synchronised with.
    """

    def __init__(self):
        object.__setattr__(self, "_h", None)
        object.__setattr__(self, "_npts", 0)

    def _require(self):
        h = object.__getattribute__(self, "_h")
        if h is None:
            raise SemicError(-1, "surface_state_class is not allocated (call surface_alloc first)")
        return h

    def _view(self, k):
        h = self._require()
        n = object.__getattribute__(self, "_npts")
        if k == "mask":
            p = lib().semic_state_mask(h)
            return np.ctypeslib.as_array(p, shape=(n,)) if n else np.zeros(0, dtype=np.int32)
        p = lib().semic_state_field(h, FIELD_ID[k])
        return np.ctypeslib.as_array(p, shape=(n,)) if n else np.zeros(0)

    def __getattr__(self, k):
        if k in FIELD_ID or k == "mask":
            return self._view(k)
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if k in FIELD_ID or k == "mask":
            self._view(k)[...] = v
        else:
            object.__setattr__(self, k, v)       # e.g. f2py/test.py:72 assigns the non-member `tt`

    # conveniences beyond the f90wrap surface
    @property
    def handle(self):
        return self._require()

    @property
    def npts(self):
        return object.__getattribute__(self, "_npts")

    def upload(self, fields=_capi.FIELDS_ALL):
        check(lib().semic_state_upload(self._require(), fields))

    def download(self, fields=_capi.FIELDS_ALL):
        check(lib().semic_state_download(self._require(), fields))

    def set_sync(self, upload_fields, download_fields):
        check(lib().semic_state_set_sync(self._require(), upload_fields, download_fields))

    def as_slab(self):
        """(31, npts) copy of the host mirror in field-id order (for comparisons)."""
        return np.stack([np.array(self._view(n)) for n in FIELD_NAMES])


class Surface_Physics_Class(object):
    """surface_physics_class (surface_physics.f90:107-119): container of par, bnd, bnd0, now, mon, ann."""

    def __init__(self):
        self.par = Surface_Param_Class()
        self.bnd = Boundary_Opt_Class()
        self.bnd0 = Boundary_Opt_Class()
        self.now = Surface_State_Class()
        self.mon = Surface_State_Class()
        self.ann = Surface_State_Class()


def fields_mask(names):
    m = 0
    for n in names:
        m |= (1 << 31) if n == "mask" else (1 << FIELD_ID[n])
    return m


# ---------------------------------------------------------------------------------------------------
# module procedures
# ---------------------------------------------------------------------------------------------------
def surface_alloc(now, npts):
    """surface_alloc(now,npts), surface_physics.f90:633-677 (contents are zero-filled here, Q8)."""
    npts = int(npts)
    h = C.c_void_p()
    check(lib().semic_surface_alloc(C.byref(h), npts))
    object.__setattr__(now, "_h", h)
    object.__setattr__(now, "_npts", npts)


def surface_dealloc(now):
    """surface_dealloc(now), surface_physics.f90:681-725."""
    h = now._require()
    check(lib().semic_surface_dealloc(h))
    object.__setattr__(now, "_h", None)
    object.__setattr__(now, "_npts", 0)


def surface_physics_par_load(par, filename):
    """surface_physics_par_load(par,filename), surface_physics.f90:729-814."""
    check(lib().semic_surface_physics_par_load(C.byref(par._c), str(filename).encode()))


def surface_boundary_define(bnd, boundary):
    """surface_boundary_define(bnd,boundary), surface_physics.f90:819-878."""
    names = list(boundary)
    buf = b"".join(padded(n) for n in names)
    check(lib().semic_surface_boundary_define(C.byref(bnd._c), buf, len(names), STRLEN))


def surface_energy_and_mass_balance(now, par, bnd, day, year):
    """surface_energy_and_mass_balance(now,par,bnd,day,year), surface_physics.f90:138-153."""
    check(lib().semic_surface_energy_and_mass_balance(now._require(), C.byref(par._c), C.byref(bnd._c), int(day), int(year)))


def energy_balance(now, par, bnd, day, year):
    """energy_balance(now,par,bnd,day,year), surface_physics.f90:158-214."""
    check(lib().semic_energy_balance(now._require(), C.byref(par._c), C.byref(bnd._c), int(day), int(year)))


def mass_balance(now, par, bnd, day, year):
    """mass_balance(now,par,bnd,day,year), surface_physics.f90:232-374."""
    check(lib().semic_mass_balance(now._require(), C.byref(par._c), C.byref(bnd._c), int(day), int(year)))


def surface_physics_average(ave, now, step, nt=None):
    """surface_physics_average(ave,now,step,nt), surface_physics.f90:565-597 (device-side).

    The averaged fields live on the device while the sums run; at step "end" the host mirror of `ave` is refreshed (as the
    Fortran shim does), so that code written against the f90wrap module can read ave.tsurf right away.
    """
    check(lib().semic_surface_physics_average(ave._require(), now._require(), str(step).encode(),
                                              -1.0 if nt is None else float(nt)))
    if str(step).strip() == "end":
        ave.download()


def print_param(par):
    """print_param(par), surface_physics.f90:882-915."""
    check(lib().semic_print_param(C.byref(par._c)))


def print_boundary_opt(bnd):
    """print_boundary_opt(bnd), surface_physics.f90:919-932."""
    check(lib().semic_print_boundary_opt(C.byref(bnd._c)))


def _arr(x):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64)))


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def sensible_heat_flux(ts, ta, wind, rhoa, csh, cap):
    """elemental sensible_heat_flux, surface_physics.f90:378-389; returns shf."""
    ts, ta, wind, rhoa = np.broadcast_arrays(_arr(ts), _arr(ta), _arr(wind), _arr(rhoa))
    ts, ta, wind, rhoa = map(np.ascontiguousarray, (ts, ta, wind, rhoa))
    out = np.empty_like(ts)
    check(lib().semic_sensible_heat_flux(_p(ts), _p(ta), _p(wind), _p(rhoa), float(csh), float(cap), _p(out), ts.size))
    return out


def latent_heat_flux(ts, wind, shum, sp, rhoatm, mask, clh, eps, cls, clv):
    """elemental latent_heat_flux, surface_physics.f90:393-430; returns (lhf, subl, evap)."""
    ts, wind, shum, sp, rhoatm = map(np.ascontiguousarray,
                                     np.broadcast_arrays(_arr(ts), _arr(wind), _arr(shum), _arr(sp), _arr(rhoatm)))
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(mask, dtype=np.int32), ts.shape))
    lhf, subl, evap = np.empty_like(ts), np.empty_like(ts), np.empty_like(ts)
    check(lib().semic_latent_heat_flux(_p(ts), _p(wind), _p(shum), _p(sp), _p(rhoatm),
                                       m.ctypes.data_as(C.POINTER(C.c_int)), float(clh), float(eps), float(cls),
                                       float(clv), _p(lhf), _p(subl), _p(evap), ts.size))
    return lhf, subl, evap


def longwave_upward(ts, sigma):
    """elemental longwave_upward, surface_physics.f90:434-438; returns lwout."""
    ts = _arr(ts)
    out = np.empty_like(ts)
    check(lib().semic_longwave_upward(_p(ts), float(sigma), _p(out), ts.size))
    return out
