"""ctypes loader for libsemic_b200.so -- the C-ABI declared in include/semic_b200.h.

There is no fallback: if the shared library is missing this module raises, and every compute entry
of the library itself fails with SEMIC_ERR_CUDA when no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SEMIC_B200_LIB: another build of the same library (A/B timing of two builds on one box, tools/ab_builds.sh); never a fallback
LIB_PATH = os.environ.get("SEMIC_B200_LIB") or os.path.join(_HERE, "lib", "libsemic_b200.so")

STRLEN = 256
NBOUNDARY = 30
NFIELDS = 31
F_MASK = 31
FIELDS_ALL = 0xFFFFFFFF
FIELDS_STATE = 0x7FFFFF          # the 23 non-forcing fields
NFORCING = 9
NOUT = 8

FIELD_NAMES = ["t2m", "tsurf", "hsnow", "hice", "alb", "alb_snow", "melt", "melted_snow", "melted_ice",
               "refr", "smb", "acc", "lhf", "shf", "lwu", "subl", "evap", "smb_snow", "smb_ice", "runoff",
               "qmr", "qmr_res", "amp", "sf", "rf", "sp", "lwd", "swd", "wind", "rhoa", "qq"]
FIELD_ID = {n: i for i, n in enumerate(FIELD_NAMES)}
FORCING_NAMES = ["sf", "rf", "swd", "lwd", "wind", "sp", "rhoa", "qq", "t2m"]      # utils.f90:63-67
OUT_NAMES = ["stt", "alb", "swnet", "smb", "melt", "acc", "shf", "lhf"]            # utils.f90:89-93
PAR_DOUBLES = ["ceff", "albi", "albl", "alb_smax", "alb_smin", "hcrit", "rcrit", "amp", "csh", "clh",
               "tmin", "tmax", "tstic", "tsticsub", "tau_a", "tau_f", "w_crit", "mcrit", "afac", "tmid"]
BND_NAMES = ["t2m", "tsurf", "hsnow", "alb", "melt", "refr", "smb", "acc", "lhf", "shf", "subl", "amp"]

B_TS_LT_T0, B_CLAMP, B_GATE, B_DIURNAL_LO, B_DIURNAL_HI, B_MELT_POS, B_HSNOW_ZERO, B_SNOW2ICE = \
    1, 2, 4, 8, 16, 32, 64, 128


class SemicPar(C.Structure):
    """semic_par == surface_param_class (surface_physics.f90:28-54)."""
    _fields_ = [("name", C.c_char * STRLEN),
                ("boundary", (C.c_char * STRLEN) * NBOUNDARY),
                ("alb_scheme", C.c_char * STRLEN),
                ("nx", C.c_int), ("n_ksub", C.c_int)] + [(n, C.c_double) for n in PAR_DOUBLES]


class SemicBnd(C.Structure):
    """semic_bnd == boundary_opt_class (surface_physics.f90:92-105)."""
    _fields_ = [(n, C.c_int) for n in BND_NAMES]


class SemicRunOpts(C.Structure):
    _fields_ = [("strict", C.c_int), ("out_days", C.c_int), ("reserved0", C.c_int), ("avg_days", C.c_int), ("reserved", C.c_int * 4)]


class SemicError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("semic_b200 error %d: %s" % (code, text))
        self.code = code


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol of include/semic_b200.h
PROTOTYPES = {
    "semic_surface_alloc": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "semic_surface_dealloc": (C.c_int, [_vp]),
    "semic_surface_physics_par_load": (C.c_int, [C.POINTER(SemicPar), C.c_char_p]),
    "semic_surface_boundary_define": (C.c_int, [C.POINTER(SemicBnd), C.c_char_p, C.c_int, C.c_int]),
    "semic_surface_energy_and_mass_balance": (C.c_int, [_vp, C.POINTER(SemicPar), C.POINTER(SemicBnd), C.c_int, C.c_int]),
    "semic_energy_balance": (C.c_int, [_vp, C.POINTER(SemicPar), C.POINTER(SemicBnd), C.c_int, C.c_int]),
    "semic_mass_balance": (C.c_int, [_vp, C.POINTER(SemicPar), C.POINTER(SemicBnd), C.c_int, C.c_int]),
    "semic_surface_physics_average": (C.c_int, [_vp, _vp, C.c_char_p, C.c_double]),
    "semic_print_param": (C.c_int, [C.POINTER(SemicPar)]),
    "semic_print_boundary_opt": (C.c_int, [C.POINTER(SemicBnd)]),
    "semic_sensible_heat_flux": (C.c_int, [_dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp, C.c_int]),
    "semic_latent_heat_flux": (C.c_int, [_dp, _dp, _dp, _dp, _dp, _ip, C.c_double, C.c_double, C.c_double,
                                         C.c_double, _dp, _dp, _dp, C.c_int]),
    "semic_longwave_upward": (C.c_int, [_dp, C.c_double, _dp, C.c_int]),
    "semic_sensible_heat_flux_1": (C.c_double, [C.c_double] * 6),
    "semic_latent_heat_flux_1": (C.c_double, [C.c_double] * 5 + [C.c_int] + [C.c_double] * 4 + [C.c_int]),
    "semic_longwave_upward_1": (C.c_double, [C.c_double] * 2),
    "semic_constant": (C.c_double, [C.c_char_p]),
    "semic_state_npts": (C.c_int, [_vp]),
    "semic_state_ld": (C.c_int, [_vp]),
    "semic_state_field": (_dp, [_vp, C.c_int]),
    "semic_state_mask": (_ip, [_vp]),
    "semic_state_device_slab": (_vp, [_vp]),
    "semic_state_upload": (C.c_int, [_vp, C.c_uint64]),
    "semic_state_download": (C.c_int, [_vp, C.c_uint64]),
    "semic_state_set_sync": (C.c_int, [_vp, C.c_uint64, C.c_uint64]),
    "semic_series_alloc": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "semic_series_free": (C.c_int, [_vp]),
    "semic_series_npts": (C.c_int, [_vp]),
    "semic_series_ld": (C.c_int, [_vp]),
    "semic_series_nvar": (C.c_int, [_vp]),
    "semic_series_ndays": (C.c_int, [_vp]),
    "semic_series_device_ptr": (_vp, [_vp]),
    "semic_series_upload": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int]),
    "semic_series_download": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int]),
    "semic_series_synth_forcing": (C.c_int, [_vp, C.c_uint64, C.c_int64, C.c_int, C.c_int]),
    "semic_state_synth_mask": (C.c_int, [_vp, C.c_uint64, C.c_int64, C.c_double, C.c_double]),
    "semic_state_fill": (C.c_int, [_vp, C.c_int, C.c_double]),
    "semic_run": (C.c_int, [_vp, C.POINTER(SemicPar), C.POINTER(SemicBnd), _vp, _vp, C.c_int, C.c_int, _vp,
                            C.POINTER(SemicRunOpts), C.POINTER(C.c_float)]),
    "semic_run_host": (C.c_int, [_vp, C.POINTER(SemicPar), C.POINTER(SemicBnd), _vp, _vp, C.c_int, C.c_int, C.c_int,
                                 _vp, C.c_int, C.c_int, C.POINTER(SemicRunOpts), C.POINTER(C.c_float)]),
    "semic_host_alloc": (_vp, [C.c_size_t]),
    "semic_host_free": (C.c_int, [_vp]),
    "semic_ensemble_cost": (C.c_int, [_vp, C.POINTER(SemicPar), C.c_int, C.POINTER(SemicBnd), _vp, _vp, C.c_int, C.c_int,
                                      _vp, _dp, C.POINTER(C.c_float)]),
    "semic_comm_unique_id": (C.c_int, [C.c_char_p]),
    "semic_comm_init": (C.c_int, [C.POINTER(_vp), C.c_char_p, C.c_int, C.c_int]),
    "semic_comm_allreduce_sum": (C.c_int, [_vp, _vp, C.c_int]),
    "semic_comm_last_allreduce_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "semic_comm_p2p": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "semic_comm_free": (C.c_int, [_vp]),
    "semic_set_device": (C.c_int, [C.c_int]),
    "semic_device_count": (C.c_int, []),
    "semic_device_synchronize": (C.c_int, []),
    "semic_last_error": (C.c_char_p, []),
    "semic_version": (C.c_char_p, []),
    "semic_kernel_launches": (C.c_int, []),
    "semic_probe_dfma": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double)]),
    "semic_probe_copy": (C.c_int, [C.c_size_t, C.POINTER(C.c_float)]),
    "semic_flush_l2": (C.c_int, []),
}


def lib():
    """Loads libsemic_b200.so (built by `make -C semic_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libsemic_b200.so is not built (%s): run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` -- there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise SemicError(rc, lib().semic_last_error().decode(errors="replace"))
    return rc


def fstr(raw):
    """Fortran-style blank padded / NUL terminated bytes -> str (trim())."""
    b = bytes(raw)
    b = b.split(b"\0", 1)[0]
    return b.decode(errors="replace").rstrip(" ")


def padded(text):
    b = text.encode() if isinstance(text, str) else bytes(text)
    b = b[:STRLEN]
    return b + b" " * (STRLEN - len(b))
