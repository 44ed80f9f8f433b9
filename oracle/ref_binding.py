"""ctypes binding of oracle/_ref/libsemic_ref.so: the reference's own Fortran sources translated to C
by oracle/f90_to_c.py and compiled here.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  Nothing under semic_b200/ does.

Everything is driven through the reference's own entry points (`surface_physics_par_load` reads a namelist
file, `surface_boundary_define` translates the names, `surface_alloc` allocates, the day loop calls
`surface_energy_and_mass_balance`); derived-type components are reached through the reflection tables the
translator emits, so this file holds no model arithmetic.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_SO = os.path.join(REF_DIR, "libsemic_ref.so")
REFERENCE = os.environ.get("SEMIC_REFERENCE", "/root/reference")

F_K_D, F_K_F, F_K_I, F_K_L, F_K_S, F_K_T = 1, 2, 3, 4, 5, 6
DESC, FIXED = 10, 20


class ArrD(C.Structure):
    _fields_ = [("a", C.POINTER(C.c_double)), ("n0", C.c_long), ("n1", C.c_long)]


class ArrI(C.Structure):
    _fields_ = [("a", C.POINTER(C.c_int)), ("n0", C.c_long), ("n1", C.c_long)]


def have_reference_tree():
    return os.path.isfile(os.path.join(REFERENCE, "surface_physics.f90"))


def available():
    """the prebuilt library travels to the GPU box; the reference tree exists only in the build container"""
    return os.path.isfile(_SO) or have_reference_tree()


def build(force=False):
    """(re)build oracle/_ref from the reference tree when it is present; otherwise use what is there"""
    if have_reference_tree():
        if force:
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "ref"])
        else:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    if not os.path.isfile(_SO):
        raise RuntimeError("oracle/_ref/libsemic_ref.so is missing and there is no reference tree to build it from")
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ref_new.argtypes = [C.c_char_p]
        L.ref_new.restype = C.c_void_p
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_sizeof.argtypes = [C.c_char_p]
        L.ref_sizeof.restype = C.c_size_t
        L.ref_comp.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_comp.restype = C.c_void_p
        L.ref_comp_name.argtypes = [C.c_char_p, C.c_int]
        L.ref_comp_name.restype = C.c_char_p
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        vp = C.c_void_p
        L.ref_run.argtypes = [vp, vp, vp, C.c_int, dp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int]
        L.ref_run.restype = C.c_int
        L.ref_surface_alloc.argtypes = [vp, ip]
        L.ref_surface_dealloc.argtypes = [vp]
        L.ref_surface_physics_par_load.argtypes = [vp, C.c_char_p, C.c_long]
        L.ref_surface_boundary_define.argtypes = [vp, C.c_void_p, C.c_long]
        L.ref_surface_energy_and_mass_balance.argtypes = [vp, vp, vp, ip, ip]
        L.ref_energy_balance.argtypes = [vp, vp, vp, ip, ip]
        L.ref_mass_balance.argtypes = [vp, vp, vp, ip, ip]
        L.ref_surface_physics_average.argtypes = [vp, vp, C.c_char_p, C.c_long, dp]
        L.ref_calculate_crmsd.argtypes = [dp, dp, ip, ip, dp]
        L.ref_sensible_heat_flux.argtypes = [dp] * 7
        L.ref_latent_heat_flux.argtypes = [dp] * 5 + [ip] + [dp] * 7
        L.ref_longwave_upward.argtypes = [dp] * 3
        L.ref_diurnal_cycle.argtypes = [dp] * 4
        L.ref_albedo_isba.argtypes = [dp] * 11
        L.ref_albedo_slater.argtypes = [dp] * 6
        L.ref_albedo_denby.argtypes = [dp] * 5
        L.ref_ew_sat.argtypes = [dp]
        L.ref_ew_sat.restype = C.c_double
        L.ref_ei_sat.argtypes = [dp]
        L.ref_ei_sat.restype = C.c_double
        for n in ("ref_surface_alloc", "ref_surface_dealloc", "ref_surface_physics_par_load", "ref_surface_boundary_define",
                  "ref_surface_energy_and_mass_balance", "ref_energy_balance", "ref_mass_balance",
                  "ref_surface_physics_average", "ref_calculate_crmsd", "ref_sensible_heat_flux", "ref_latent_heat_flux",
                  "ref_longwave_upward", "ref_diurnal_cycle", "ref_albedo_isba", "ref_albedo_slater", "ref_albedo_denby"):
            getattr(L, n).restype = None
        _lib = L
    return _lib


def _d(x):
    return C.byref(C.c_double(x))


class Obj:
    """an instance of one of the reference's derived types, components by name"""

    def __init__(self, tname):
        self.__dict__["_t"] = tname.encode()
        self.__dict__["_p"] = lib().ref_new(self._t)
        if not self._p:
            raise KeyError(tname)

    def _comp(self, name):
        kind, clen = C.c_int(), C.c_int()
        addr = lib().ref_comp(self._t, self._p, name.encode(), C.byref(kind), C.byref(clen))
        if not addr:
            raise AttributeError(name)
        return addr, kind.value, clen.value

    def components(self):
        out, i = [], 0
        while True:
            n = lib().ref_comp_name(self._t, i)
            if n is None:
                return out
            out.append(n.decode())
            i += 1

    def __getattr__(self, name):
        addr, kind, clen = self._comp(name)
        if kind == F_K_D:
            return C.c_double.from_address(addr).value
        if kind in (F_K_I, F_K_L):
            return C.c_int.from_address(addr).value
        if kind == F_K_S:
            return C.string_at(addr, clen).decode().rstrip(" ")
        if kind == F_K_D + DESC:
            d = ArrD.from_address(addr)
            n = d.n0 * max(d.n1, 1)
            return np.ctypeslib.as_array(d.a, shape=(n,)) if n else np.zeros(0)
        if kind == F_K_I + DESC:
            d = ArrI.from_address(addr)
            return np.ctypeslib.as_array(d.a, shape=(d.n0,)) if d.n0 else np.zeros(0, dtype=np.int32)
        raise AttributeError("%s: kind %d not exposed" % (name, kind))

    def __setattr__(self, name, value):
        addr, kind, clen = self._comp(name)
        if kind == F_K_D:
            C.c_double.from_address(addr).value = float(value)
        elif kind in (F_K_I, F_K_L):
            C.c_int.from_address(addr).value = int(value)
        elif kind == F_K_S:
            b = str(value).encode()[:clen].ljust(clen, b" ")
            C.memmove(addr, b, clen)
        elif kind in (F_K_D + DESC, F_K_I + DESC):
            getattr(self, name)[:] = value
        else:
            raise AttributeError(name)

    def addr(self, name):
        return self._comp(name)[0]

    def free(self):
        if self._p:
            lib().ref_free(self._p)
            self.__dict__["_p"] = None


def _nml_value(v):
    if isinstance(v, str):
        return '"%s"' % v
    if isinstance(v, (int, np.integer)):
        return "%d" % v
    r = repr(float(v))
    return r.replace("e", "d") if "e" in r else r


def write_namelist(path, par_d, boundary=()):
    """a &surface_physics group in the reference's own file format (example/example.namelist:10-31)"""
    names = list(boundary) + [""] * max(0, 3 - len(boundary))
    lines = ["&surface_physics", "  boundary = " + " ".join('"%s"' % b for b in names) + ","]
    for k in sorted(par_d):
        if k in ("nx", "name", "tsticsub"):
            continue
        lines.append("  %s = %s," % (k, _nml_value(par_d[k])))
    with open(path, "w") as f:
        f.write("\n".join(lines + ["/", ""]))
    return path


def load_par(par_d, boundary=(), nx=0):
    """par through the reference's own `surface_physics_par_load` (surface_physics.f90:729-814) and bnd through
    `surface_boundary_define` (:819-878).  Incoming defaults are zero / blank (Q10)."""
    par, bnd = Obj("surface_param_class"), Obj("boundary_opt_class")
    par.nx = nx
    fd, path = tempfile.mkstemp(suffix=".nml")
    os.close(fd)
    try:
        write_namelist(path, par_d, boundary)
        b = path.encode()
        lib().ref_surface_physics_par_load(par._p, b, len(b))
    finally:
        os.unlink(path)
    lib().ref_surface_boundary_define(bnd._p, par.addr("boundary"), 30)
    return par, bnd


FIELDS = ["t2m", "tsurf", "hsnow", "hice", "alb", "alb_snow", "melt", "melted_snow", "melted_ice",
          "refr", "smb", "acc", "lhf", "shf", "lwu", "subl", "evap", "smb_snow", "smb_ice", "runoff",
          "qmr", "qmr_res", "amp", "sf", "rf", "sp", "lwd", "swd", "wind", "rhoa", "qq"]


def new_state(nx):
    now = Obj("surface_state_class")
    lib().ref_surface_alloc(now._p, C.byref(C.c_int(nx)))
    return now


def state_from_slab(slab, mask, nx=None):
    nx = slab.shape[1] if nx is None else nx
    now = new_state(nx)
    for i, n in enumerate(FIELDS):
        getattr(now, n)[:] = slab[i, :nx]
    now.mask[:] = mask[:nx]
    return now


def state_to_slab(now, slab):
    nx = now.mask.shape[0]
    for i, n in enumerate(FIELDS):
        slab[i, :nx] = getattr(now, n)


def free_state(now):
    lib().ref_surface_dealloc(now._p)
    now.free()


def run(slab, mask, par_d, bnd_names, forcing, ndays, vali=None, out_days=0, nx=None):
    """same contract as oracle.binding.run (without the branch bitmap): the slab is updated in place,
    returns out [out_days][8][nx] or None"""
    nx = slab.shape[1] if nx is None else nx
    nwin, nine, ldf = forcing.shape
    assert nine == 9 and forcing.flags.c_contiguous and forcing.dtype == np.float64
    par, bnd = load_par(par_d, bnd_names, nx)
    now = state_from_slab(slab, mask, nx)
    out = np.zeros((out_days, 8, nx)) if out_days > 0 else None
    dp = C.POINTER(C.c_double)
    rc = lib().ref_run(now._p, par._p, bnd._p, nx, forcing.ctypes.data_as(dp),
                       vali.ctypes.data_as(dp) if vali is not None else None, nwin, ldf, ndays,
                       out.ctypes.data_as(dp) if out is not None else None, max(out_days, 1), nx, out_days)
    assert rc == 0
    state_to_slab(now, slab)
    free_state(now)
    par.free()
    bnd.free()
    return out


def calculate_crmsd(x1, x2):
    """x1, x2: [ntime][nx] (= Fortran (nx,ntime) column-major), utils.f90:23-42"""
    ntime, nx = x1.shape
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    r = np.zeros(nx)
    dp = C.POINTER(C.c_double)
    lib().ref_calculate_crmsd(x1.ctypes.data_as(dp), x2.ctypes.data_as(dp), C.byref(C.c_int(ntime)),
                              C.byref(C.c_int(nx)), r.ctypes.data_as(dp))
    return r


def run_program(name, args, cwd):
    """run one of the translated reference programs (`example`, `run_particles`) in `cwd`"""
    build()
    exe = os.path.join(REF_DIR, name + "_ref.x")
    return subprocess.run([exe] + list(args), cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, check=True).stdout
