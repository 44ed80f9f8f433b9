"""Device-resident time series and the multi-day / host-streamed / ensemble entry points of the C-ABI."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import NFORCING, NOUT, SemicRunOpts, check, lib


class Series(object):
    """semic_series: device array [day][var][point] (point-contiguous, padded leading dimension)."""

    def __init__(self, npts, nvar, ndays):
        self._h = C.c_void_p()
        check(lib().semic_series_alloc(C.byref(self._h), int(npts), int(nvar), int(ndays)))
        self.npts, self.nvar, self.ndays = int(npts), int(nvar), int(ndays)

    @property
    def handle(self):
        return self._h

    @property
    def ld(self):
        return lib().semic_series_ld(self._h)

    def upload(self, host, day0=0):
        """host: float64 [ndays][nvar][npts]"""
        host = np.ascontiguousarray(host, dtype=np.float64)
        nd, nv, npts = host.shape
        assert nv == self.nvar and npts >= self.npts
        check(lib().semic_series_upload(self._h, int(day0), nd, host.ctypes.data_as(C.POINTER(C.c_double)), npts))
        return self

    def download(self, day0=0, ndays=None, point0=0, npoints=None):
        ndays = self.ndays - day0 if ndays is None else ndays
        npoints = self.npts - point0 if npoints is None else npoints
        out = np.empty((ndays, self.nvar, npoints), dtype=np.float64)
        check(lib().semic_series_download(self._h, int(day0), int(ndays), out.ctypes.data_as(C.POINTER(C.c_double)),
                                          int(npoints), int(point0), int(npoints)))
        return out

    def synth_forcing(self, seed, point0=0, dayofs=0, daystride=1):
        check(lib().semic_series_synth_forcing(self._h, int(seed), int(point0), int(dayofs), int(daystride)))
        return self

    def free(self):
        if self._h:
            check(lib().semic_series_free(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def run(state, par, bnd, forcing, ndays, vali=None, day0=0, out=None, out_days=0, strict=False, avg_days=0):
    """semic_run: ndays days of the persistent kernel on device-resident forcing.  Returns kernel ms.
    avg_days > 1: rows of `out` are means over consecutive periods of avg_days days (in-kernel surface_physics_average)."""
    opts = SemicRunOpts()
    opts.strict = 1 if strict else 0
    opts.out_days = int(out_days)
    opts.avg_days = int(avg_days)
    ms = C.c_float(0.0)
    check(lib().semic_run(state.handle, C.byref(par._c), C.byref(bnd._c), forcing.handle,
                          vali.handle if vali is not None else None, int(day0), int(ndays),
                          out.handle if out is not None else None, C.byref(opts), C.byref(ms)))
    return ms.value


class PinnedArray(object):
    """float64 numpy array in pinned host memory (semic_host_alloc)."""

    def __init__(self, shape):
        n = int(np.prod(shape))
        self._p = lib().semic_host_alloc(max(n, 1) * 8)
        if not self._p:
            raise _capi.SemicError(-2, lib().semic_last_error().decode())
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_double)), shape=(max(n, 1),))[:n].reshape(shape)

    @property
    def ptr(self):
        return self._p

    def free(self):
        if self._p:
            self.array = None
            lib().semic_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def run_host(state, par, bnd, h_forcing, ndays, h_vali=None, h_out=None, out_days=0, chunk_days=8, strict=False, avg_days=0):
    """semic_run_host: forcing [nwin][9][ldh] (and vali [nwin][8][ldh]) live in HOST memory (numpy arrays or
    PinnedArray.array); H2D, kernel and D2H are pipelined in chunks of chunk_days.  Returns total ms.
    avg_days > 1: h_out holds ceil(out_days/avg_days) rows of period means instead of out_days daily rows."""
    nwin, nine, ldh = h_forcing.shape
    assert nine == NFORCING and h_forcing.flags.c_contiguous and h_forcing.dtype == np.float64
    if h_vali is not None:
        assert h_vali.shape == (nwin, NOUT, ldh) and h_vali.flags.c_contiguous
    if out_days > 0:
        rows = out_days if avg_days <= 1 else (out_days + avg_days - 1) // avg_days
        assert h_out is not None and h_out.shape == (rows, NOUT, ldh) and h_out.flags.c_contiguous
    opts = SemicRunOpts()
    opts.strict = 1 if strict else 0
    opts.avg_days = int(avg_days)
    ms = C.c_float(0.0)
    check(lib().semic_run_host(state.handle, C.byref(par._c), C.byref(bnd._c), h_forcing.ctypes.data,
                               h_vali.ctypes.data if h_vali is not None else None, nwin, ldh, int(ndays),
                               h_out.ctypes.data if h_out is not None else None, int(out_days), int(chunk_days),
                               C.byref(opts), C.byref(ms)))
    return ms.value


def ensemble_cost(init_state, pars, bnd, forcing, vali, ntime, nloop, comm=None):
    """semic_ensemble_cost: per-particle sum of squared CRMSDs over the local points (all-reduced over `comm`).

    pars: list of Surface_Param_Class.  Returns (sumsq float64[npar], kernel ms)."""
    npar = len(pars)
    arr = (_capi.SemicPar * npar)()
    for k, p in enumerate(pars):
        C.memmove(C.addressof(arr) + k * C.sizeof(_capi.SemicPar), C.addressof(p._c), C.sizeof(_capi.SemicPar))
    sumsq = np.zeros(npar, dtype=np.float64)
    ms = C.c_float(0.0)
    check(lib().semic_ensemble_cost(init_state.handle, arr, npar, C.byref(bnd._c), forcing.handle, vali.handle,
                                    int(ntime), int(nloop), comm, sumsq.ctypes.data_as(C.POINTER(C.c_double)),
                                    C.byref(ms)))
    return sumsq, ms.value


def set_device(dev):
    check(lib().semic_set_device(int(dev)))


def device_count():
    return lib().semic_device_count()


def kernel_launches():
    return lib().semic_kernel_launches()
