"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/semic_b200.h declares,
the host-only entries (namelist, boundary flags, printing) behave like the reference, and compute entries fail
loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from semic_b200 import _capi, driver, synth
from semic_b200 import surface_physics as sp
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "semic_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(semic_[a-z0-9_]+)\s*\(", hdr)))


def _have_gpu():
    try:
        return _capi.lib().semic_device_count() > 0
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 50
    L = _capi.lib()
    nm = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (semic_[a-z0-9_]+)", nm))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    # the ctypes prototype table binds the same set
    assert set(names) == set(_capi.PROTOTYPES), set(names) ^ set(_capi.PROTOTYPES)
    assert b"sm_100a" in L.semic_version()


def test_library_does_not_link_the_oracle_or_torch():
    ldd = subprocess.check_output(["ldd", _capi.LIB_PATH], text=True)
    assert "oracle" not in ldd and "torch" not in ldd and "libc10" not in ldd
    nm = subprocess.check_output(["nm", "-D", _capi.LIB_PATH], text=True)
    assert "oracle_" not in nm


def test_struct_layout_matches_header():
    # semic_par: 256 + 30*256 + 256 chars, 2 ints, 20 doubles; semic_bnd: 12 ints
    assert C.sizeof(_capi.SemicPar) == 256 * 32 + 8 + 20 * 8
    assert C.sizeof(_capi.SemicBnd) == 48
    assert _capi.SemicPar.ceff.offset == 256 * 32 + 8


def test_par_load_example_namelist(tmp_path):
    par = sp.Surface_Param_Class()
    par.afac = 123.0                      # not in the file: keeps the incoming value (surface_physics.f90:752-775)
    par.tmid = -4.0
    sp.surface_physics_par_load(par, util.write_example_namelist(str(tmp_path / "example.namelist")))
    exp = util.PAR_EXAMPLE
    for k in ("tstic", "ceff", "csh", "clh", "alb_smax", "alb_smin", "albi", "albl", "tmin", "tmax", "hcrit", "rcrit",
              "amp", "tau_a", "tau_f", "w_crit", "mcrit"):
        assert getattr(par, k) == exp[k], k
    assert par.n_ksub == 3 and par.alb_scheme == "none"
    assert par.tsticsub == 86400.0 / 3.0                                   # :810
    assert par.afac == 123.0 and par.tmid == -4.0
    assert par.boundary[:3] == ["", "", ""] and all(b == "" for b in par.boundary)


def test_par_load_particle_namelist_and_foreign_groups():
    par = sp.Surface_Param_Class()
    par.alb_smin = 0.5
    sp.surface_physics_par_load(par, os.path.join(util.GOLDEN, "particle.nml"))      # comma separated strings, &smb_output after
    assert par.hcrit == 0.25 and par.alb_smax == 0.85 and par.amp == 2.5 and par.rcrit == 0.5
    assert par.albi == 0.45 and par.albl == 0.225 and par.afac == -0.18 and par.tmid == 275.35
    assert par.alb_scheme == "none" and par.n_ksub == 3 and par.alb_smin == 0.5


def test_par_load_syntax_variants(tmp_path):
    f = tmp_path / "v.nml"
    f.write_text("""! leading comment
&other  x = 1, s = 'a / b', /
$SURFACE_PHYSICS
  BOUNDARY(2:3) = 'tsurf', "alb"   ! section
  boundary(5) = 'smb'
  TSTIC = 8.64D4 , ceff=2d6
  csh = 2.0E-3, clh = 5.0e-4,
  alb_scheme = 'isba'
  n_ksub = 24
  boundary(7:9) = 3*'amp'
$END
""")
    par = sp.Surface_Param_Class()
    sp.surface_physics_par_load(par, str(f))
    b = par.boundary
    assert b[0] == "" and b[1] == "tsurf" and b[2] == "alb" and b[4] == "smb" and b[6:9] == ["amp"] * 3
    assert par.tstic == 86400.0 and par.ceff == 2e6 and par.csh == 2e-3 and par.alb_scheme == "isba" and par.n_ksub == 24
    assert par.tsticsub == 3600.0
    bnd = sp.Boundary_Opt_Class()
    sp.surface_boundary_define(bnd, par.boundary)
    assert bnd.tsurf and bnd.alb and bnd.smb and bnd.amp and not bnd.melt and not bnd.t2m


@pytest.mark.parametrize("text,code", [
    ("&driver\n nx = 1\n/\n", -4),                                          # group missing
    ("&surface_physics\n bogus = 1\n/\n", -4),                              # unknown member
    ("&surface_physics\n n_ksub = 3.5\n/\n", -4),                           # bad integer
    ("&surface_physics\n ceff = abc\n/\n", -4),                             # bad real
    ("&surface_physics\n ceff = 1.0\n", -4),                                # no terminator
])
def test_par_load_errors(tmp_path, text, code):
    f = tmp_path / "bad.nml"
    f.write_text(text)
    par = sp.Surface_Param_Class()
    with pytest.raises(_capi.SemicError) as e:
        sp.surface_physics_par_load(par, str(f))
    assert e.value.code == code
    with pytest.raises(_capi.SemicError) as e:
        sp.surface_physics_par_load(par, str(tmp_path / "nope.nml"))
    assert e.value.code == -3


def test_boundary_define_names():
    bnd = sp.Boundary_Opt_Class()
    sp.surface_boundary_define(bnd, ["t2m", "tsurf", "hsnow", "alb", "melt", "refr", "smb", "acc", "lhf", "shf", "subl", "amp"])
    assert all(getattr(bnd, n) for n in _capi.BND_NAMES)
    sp.surface_boundary_define(bnd, ["massbal", "TSURF", " alb", "alb "])     # Q11: unknown names are ignored; trim() only
    assert not bnd.smb and not bnd.tsurf and bnd.alb
    assert sum(getattr(bnd, n) for n in _capi.BND_NAMES) == 1
    sp.surface_boundary_define(bnd, [])
    assert not any(getattr(bnd, n) for n in _capi.BND_NAMES)


def test_print_param_and_boundary_opt(capfd, tmp_path):
    par = sp.Surface_Param_Class()
    sp.surface_physics_par_load(par, util.write_slater_namelist(str(tmp_path / "test.nml")))
    par.nx = 7
    par.boundary = ["tsurf"]
    sp.print_param(par)
    bnd = sp.Boundary_Opt_Class()
    sp.surface_boundary_define(bnd, par.boundary)
    sp.print_boundary_opt(bnd)
    out = capfd.readouterr().out.splitlines()
    assert out[0] == "boundary tsurf" and out[1] == "alb_scheme  slater"
    assert out[2] == "nx             7" and out[3] == "n_ksub         3"
    assert "ceff        0.200000E+07" in out and "tstic        86400.0    " in out
    assert "albi        0.450000    " in out and "mcrit       0.600000E-07" in out
    assert "tsurf   T" in out and "alb     F" in out and "amp     F" in out


def test_constants():
    L = _capi.lib()
    for n, v in (("sigm", 5.67e-8), ("cap", 1000.0), ("eps", 0.62197), ("cls", 2.83e6), ("clv", 2.5e6)):
        assert L.semic_constant(n.encode()) == v == getattr(sp, n)
    assert np.isnan(L.semic_constant(b"nope"))


def test_compute_entries_fail_loudly_without_a_gpu():
    if _capi.lib().semic_device_count() > 0:
        pytest.skip("a GPU is present")
    st = sp.Surface_State_Class()
    with pytest.raises(_capi.SemicError) as e:
        sp.surface_alloc(st, 16)
    assert e.value.code == -2 and "CUDA" in str(e.value)
    with pytest.raises(_capi.SemicError):
        sp.longwave_upward(np.array([273.15]), sp.sigm)
    # the one-value forms behind the Fortran shim's ELEMENTAL specifics return NaN and leave the error text
    L = _capi.lib()
    assert np.isnan(L.semic_longwave_upward_1(273.15, sp.sigm))
    assert np.isnan(L.semic_sensible_heat_flux_1(260.0, 262.0, 6.0, 1.2, 2e-3, sp.cap))
    assert np.isnan(L.semic_latent_heat_flux_1(260.0, 6.0, 1e-3, 9e4, 1.2, 2, 5e-4, sp.eps, sp.cls, sp.clv, 0))
    assert b"CUDA" in L.semic_last_error()


def test_driver_namelist_and_shards(tmp_path):
    d = driver.read_driver_namelist(util.write_example_namelist(str(tmp_path / "example.namelist")))
    assert (d["nloop"], d["ntime"], d["nx"]) == (100, 365, 7)
    assert list(d["loi_mask"][:7]) == [1, 1, 1, 2, 2, 2, 2] and d["output"] == "example.out"
    for n, w in ((16777216, 8), (100000, 8), (7, 3), (5, 8), (0, 2)):
        edges = [driver.shard_bounds(n, w, r) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[r][1] == edges[r + 1][0] for r in range(w - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        driver.shard_bounds(10, 2, 2)


def test_ascii_readers_roundtrip(tmp_path):
    forc, vali = util.load_fixture("transect")
    ff, vf = tmp_path / "f.txt", tmp_path / "v.txt"
    np.savetxt(ff, forc.reshape(365, -1)[:20])
    np.savetxt(vf, vali.reshape(365, -1)[:20])
    assert np.array_equal(driver.read_forcing(str(ff), 20, 7), forc[:20])
    assert np.array_equal(driver.read_validation(str(vf), 20, 7), vali[:20])
    with pytest.raises(ValueError):
        driver.read_forcing(str(ff), 20, 6)


def test_numpy_forcing_generator_ranges():
    f = synth.forcing(2000, 365)
    sf, rf, swd, lwd, wind, spr, rhoa, qq, t2m = [f[:, v, :] for v in range(9)]
    assert np.isfinite(f).all() and sf.min() >= 0 and rf.min() >= 0
    assert swd.min() == 0 and 300 < swd.max() <= 400 and lwd.min() >= 70 and lwd.max() <= 340
    assert wind.min() >= 0.1 and wind.max() <= 22 and 200 < t2m.min() and t2m.max() < 300
    assert np.array_equal(synth.forcing(100, 10, point0=500, dayofs=50), synth.forcing(700, 70)[50:60, :, 500:600])
    assert np.array_equal(synth.forcing(50, 73, daystride=5)[3], synth.forcing(50, 16)[15])
    m = synth.mask(200000)
    frac = np.bincount(m, minlength=3) / m.size
    assert abs(frac[2] - 0.85) < 0.01 and abs(frac[1] - 0.10) < 0.01


# ------------------------------------------------------------------------------------------------------------
# the native command-line drivers (semic_b200/csrc/drivers): host logic runs anywhere, compute fails loudly without a GPU
# ------------------------------------------------------------------------------------------------------------
def _example_dir(tmp_path, nloop=2):
    from tests import util
    forc, vali = util.load_fixture("transect")
    (tmp_path / "data").mkdir()
    np.savetxt(tmp_path / "data" / "transect_input.txt", forc.reshape(365, -1))
    np.savetxt(tmp_path / "data" / "transect_output.txt", vali.reshape(365, -1))
    util.write_example_namelist(str(tmp_path / "example.namelist"), nloop=nloop)
    return tmp_path


def test_example_x_host_side_and_loud_failure_without_gpu(tmp_path):
    import subprocess
    exe = os.path.join(ROOT, "semic_b200", "bin", "example.x")
    assert os.path.exists(exe), "run `make -C semic_b200/csrc` (or __graft_entry__.build())"
    d = _example_dir(tmp_path)
    r = subprocess.run([exe, "example.namelist"], cwd=str(d), capture_output=True, text=True)
    # namelist, both ASCII tables and print_param are host side (example.f90:52-68)
    assert "Read forcing from file" in r.stdout and "Read validation from file" in r.stdout
    assert "alb_smax" in r.stdout and "0.810000" in r.stdout
    if not _have_gpu():
        assert r.returncode != 0 and "CUDA" in r.stderr and not (d / "example.out").exists()
    r = subprocess.run([exe], cwd=str(d), capture_output=True, text=True)                        # example.f90:38-42
    assert r.returncode != 0 and "Namelist file is missing" in r.stderr
    r = subprocess.run([exe, "nope.namelist"], cwd=str(d), capture_output=True, text=True)       # example.f90:44-48
    assert r.returncode != 0 and "not found" in r.stderr


def test_run_particles_x_stops_at_the_first_missing_namelist(tmp_path):
    import subprocess
    exe = os.path.join(ROOT, "semic_b200", "bin", "run_particles.x")
    assert os.path.exists(exe)
    d = _example_dir(tmp_path)
    (d / "optimization.namelist").write_text((d / "example.namelist").read_text())
    r = subprocess.run([exe, '"p_"', "0", "3"], cwd=str(d), capture_output=True, text=True)      # no p_000000.nml: nothing to do
    assert r.returncode == 0 and "p_000000.nml' not found, exiting loop" in r.stdout              # run_particles.f90:110-113
    assert "end particle calculation" in r.stdout


def test_binary_series_roundtrip_and_ascii_conversion(tmp_path):
    from semic_b200 import io as sio
    forc, vali = util.load_fixture("transect")
    sio.write_series(str(tmp_path / "f.bin"), forc)
    m = sio.open_series(str(tmp_path / "f.bin"))
    assert m.shape == forc.shape and np.array_equal(np.asarray(m), forc)
    np.savetxt(tmp_path / "v.txt", vali.reshape(365, -1))                    # the reference's ASCII layout, utils.f90:89-93
    sio.ascii_to_binary(str(tmp_path / "v.txt"), str(tmp_path / "v.bin"), 365, 7, nvar=8)
    assert np.array_equal(np.asarray(sio.open_series(str(tmp_path / "v.bin"))), vali)
    w = sio.create_series(str(tmp_path / "o.bin"), 3, 8, 5)
    w[...] = 1.5
    w.flush()
    assert np.array_equal(np.asarray(sio.open_series(str(tmp_path / "o.bin"))), np.full((3, 8, 5), 1.5))
    (tmp_path / "bad.bin").write_bytes(b"x" * 100)
    with pytest.raises(ValueError):
        sio.open_series(str(tmp_path / "bad.bin"))


def test_namelist_accepts_underflowing_reals_and_rejects_overflow(tmp_path):
    """a value that underflows to a subnormal is what Fortran's list-directed input stores too; overflow stays an error"""
    from semic_b200 import surface_physics as sp
    from semic_b200._capi import SemicError
    p = tmp_path / "u.nml"
    p.write_text("&surface_physics\n mcrit = 1.0d-320,\n hcrit = 2.5e-3,\n/\n")
    par = sp.Surface_Param_Class()
    sp.surface_physics_par_load(par, str(p))
    assert par.mcrit == 1e-320 and par.hcrit == 2.5e-3
    p.write_text("&surface_physics\n mcrit = 1.0d400,\n/\n")
    with pytest.raises(SemicError):
        sp.surface_physics_par_load(par, str(p))
