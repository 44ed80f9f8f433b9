"""Day loops of the reference's driver programs on top of the CUDA engine.

  run_example      example/example.f90      (config 1: &driver namelist, nloop x ntime days, writes example.out)
  run_particles    optimize/run_particles.f90 + optimize/tools.py:46-63 (config 5: parameter ensemble -> cost dict)
  shard_bounds     contiguous point ranges per rank (no communication between shards)

The numerics are the library's (libsemic_b200.so); this module is host logic only.
"""
import os
import re

import numpy as np

from . import engine
from . import surface_physics as sp
from ._capi import NFORCING, NOUT


# ---------------------------------------------------------------------------------------------------------------
# sharding: grid points are independent, each rank owns one contiguous range
# ---------------------------------------------------------------------------------------------------------------
def shard_bounds(npts, world, rank):
    """[lo, hi) of rank's contiguous point range; ranges differ by at most one point and cover [0, npts)."""
    if world < 1 or not (0 <= rank < world) or npts < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(npts, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ---------------------------------------------------------------------------------------------------------------
# file formats of the reference (utils.f90:44-95): whitespace separated ASCII, one line per day
# ---------------------------------------------------------------------------------------------------------------
def read_forcing(fnm, ntime, nx):
    """utils.read_forcing: 9*nx columns per line in the order sf rf swd lwd wind sp rhoa qq t2m -> [ntime][9][nx]"""
    a = np.loadtxt(fnm, ndmin=2)
    if a.shape[0] < ntime or a.shape[1] != NFORCING * nx:
        raise ValueError("%s: expected >= %d lines of %d columns, found %s" % (fnm, ntime, NFORCING * nx, a.shape))
    return np.ascontiguousarray(a[:ntime].reshape(ntime, NFORCING, nx))


def read_validation(fnm, ntime, nx):
    """utils.read_validation: 8*nx columns per line in the order stt alb swnet smb melt acc shf lhf -> [ntime][8][nx]"""
    a = np.loadtxt(fnm, ndmin=2)
    if a.shape[0] < ntime or a.shape[1] != NOUT * nx:
        raise ValueError("%s: expected >= %d lines of %d columns, found %s" % (fnm, ntime, NOUT * nx, a.shape))
    return np.ascontiguousarray(a[:ntime].reshape(ntime, NOUT, nx))


def read_driver_namelist(path):
    """&driver group of example.namelist / optimization.namelist (example.f90:26): nloop, ntime, nx, loi_mask(100),
    input_forcing, output, validation.  loi_mask defaults to 2 everywhere (example.f90:51)."""
    text = open(path).read()
    m = re.search(r"&driver(.*?)^\s*/", text, re.S | re.M | re.I)
    if not m:
        raise ValueError("group &driver not found in %s" % path)
    body = re.sub(r"!.*", "", m.group(1))
    out = {"loi_mask": np.full(100, 2, dtype=np.int32), "output": "", "validation": "", "input_forcing": ""}
    for key, sub, val in re.findall(r"(\w+)\s*(\([^)]*\))?\s*=\s*((?:\"[^\"]*\"|'[^']*'|[^=])*?)(?=\s*\w+\s*(?:\([^)]*\))?\s*=|\s*$)",
                                    body, re.S):
        key = key.lower()
        val = val.strip().rstrip(",").strip()
        if key in ("nloop", "ntime", "nx"):
            out[key] = int(val)
        elif key in ("input_forcing", "output", "validation"):
            out[key] = val.strip("\"'")
        elif key == "loi_mask":
            vals = [int(float(v)) for v in re.split(r"[,\s]+", val) if v]
            lo = 1
            if sub:
                s = sub.strip("()").split(":")
                lo = int(s[0]) if s[0].strip() else 1
            out["loi_mask"][lo - 1:lo - 1 + len(vals)] = vals
    return out


# ---------------------------------------------------------------------------------------------------------------
# example.f90
# ---------------------------------------------------------------------------------------------------------------
def run_example(nml_file, workdir=None, write_output=True, strict=False):
    """The program of example/example.f90 on the GPU.  Returns (out [ntime][8][nx], state, par, bnd)."""
    workdir = workdir or os.path.dirname(os.path.abspath(nml_file))
    drv = read_driver_namelist(nml_file)
    nx, ntime, nloop = drv["nx"], drv["ntime"], drv["nloop"]
    rel = lambda p: p if os.path.isabs(p) else os.path.join(workdir, p)
    forc = read_forcing(rel(drv["input_forcing"]), ntime, nx)                       # example.f90:60
    vali = read_validation(rel(drv["validation"]), ntime, nx)                       # example.f90:61
    par = sp.Surface_Param_Class()
    par.nx = nx
    sp.surface_physics_par_load(par, nml_file)                                      # example.f90:67
    now = sp.Surface_State_Class()
    sp.surface_alloc(now, nx)                                                       # example.f90:71
    now.mask = drv["loi_mask"][:nx]                                                 # example.f90:74-79
    now.hsnow = 5.0
    now.hice = 0.0
    now.alb = np.float64(np.float32(0.8))        # Q9: the reference assigns the single-precision literal 0.8
    now.tsurf = 260.0
    now.alb_snow = np.float64(np.float32(0.8))
    now.upload()
    bnd = sp.Boundary_Opt_Class()
    sp.surface_boundary_define(bnd, par.boundary)                                   # example.f90:82
    fs = engine.Series(nx, NFORCING, ntime).upload(forc)
    vs = engine.Series(nx, NOUT, ntime).upload(vali)
    outs = engine.Series(nx, NOUT, ntime)
    engine.run(now, par, bnd, fs, nloop * ntime, vali=vs, out=outs, out_days=ntime, strict=strict)   # example.f90:88-127
    now.download()
    out = outs.download()
    if write_output and drv["output"]:
        with open(rel(drv["output"]), "w") as f:                                    # example.f90:86,117-122
            f.write(" # tsurf(K)      alb      hsnow(m)      smb(m/s)     melt(m/s)      acc(m/s)\n")
            for d in range(ntime):
                f.write(" ".join("%24.16E" % v for v in out[d].reshape(-1)) + "\n")
    for s in (fs, vs, outs):
        s.free()
    return out, now, par, bnd


# ---------------------------------------------------------------------------------------------------------------
# run_particles.f90 + tools.run_particles
# ---------------------------------------------------------------------------------------------------------------
# what optimize/tools.py:7-43 writes into every particle's namelist
PARTICLE_FIXED = dict(tstic=86400., ceff=2.0e6, csh=2.0e-3, clh=5.0e-4, tmin=263.15, tmax=273.15, afac=-0.18,
                      tmid=275.35, n_ksub=3, alb_scheme="none")


def particle_params(names, pop, base=None):
    """pop: list of dicts with "id" and "pos" (optimize/tools.py); names: tuned parameter names (namelist.ranges).
    Values go through the same '%.8f' text round trip as tools.write_namelist (tools.py:35)."""
    pars = []
    for p in pop:
        par = sp.Surface_Param_Class()
        for k, v in dict(base or PARTICLE_FIXED).items():
            setattr(par, k, v)
        for n, v in zip(names, p["pos"]):
            setattr(par, n, float("%.8f" % v))
        par.tsticsub = par.tstic / float(par.n_ksub)                                 # surface_physics.f90:810
        pars.append(par)
    return pars


def finish_costs(pop, sumsq):
    """cost = sqrt(sum of squared CRMSDs) (run_particles.f90:175); NaN -> finfo.max (tools.py:62)."""
    cost = {}
    for p, s in zip(pop, sumsq):
        c = float(np.sqrt(s))
        cost[p["id"]] = np.finfo("d").max if np.isnan(c) else c
    return cost


def run_particles(pop, names, forcing, vali, mask, init, ntime, nloop, comm=None, allreduce=None, partial_fn=None):
    """In-process replacement of tools.run_particles (namelist files + mpiexec + .out files).

    forcing [ntime][9][nx_local], vali [ntime][8][nx_local], mask [nx_local]: this rank's shard of the grid;
    init: dict of initial fields (run_particles.f90:77-82: hsnow, hice, alb, tsurf, alb_snow).
    The per-particle sums of squared CRMSDs of all shards are added with ONE all-reduce: `comm` (NCCL, inside the
    library) or `allreduce(array)` (e.g. torch.distributed on gloo in CPU tests).  `partial_fn` replaces the GPU
    evaluation of the local partial sums (tests inject the oracle there).  Returns {id: cost}."""
    pars = particle_params(names, pop)
    if partial_fn is not None:
        sumsq = np.asarray(partial_fn(pars, forcing, vali, mask, init, ntime, nloop), dtype=np.float64)
    else:
        nx = forcing.shape[2]
        st = sp.Surface_State_Class()
        sp.surface_alloc(st, nx)
        st.mask = mask
        for k, v in init.items():
            setattr(st, k, v)
        st.upload()
        fs = engine.Series(nx, NFORCING, ntime).upload(forcing)
        vs = engine.Series(nx, NOUT, ntime).upload(vali)
        bnd = sp.Boundary_Opt_Class()
        sumsq, _ = engine.ensemble_cost(st, pars, bnd, fs, vs, ntime, nloop, comm=comm)
        fs.free()
        vs.free()
        sp.surface_dealloc(st)
    if allreduce is not None:
        sumsq = allreduce(sumsq)
    return finish_costs(pop, sumsq)
