"""Pins the hand-written oracle (oracle/semic_oracle.c) to the REFERENCE ITSELF.

oracle/_ref is built by `make -C oracle ref`: the reference's own surface_physics.f90, utils.f90, example.f90 and
run_particles.f90, translated statement by statement by the physics-free translator oracle/f90_to_c.py and compiled
with the oracle's flags (there is no Fortran compiler in the image).  These tests demand BIT-FOR-BIT equality between
that build and the restatement, through the reference's own entry points (par_load from a namelist file,
boundary_define, alloc, the day loop), and between the translated `example` / `run_particles` PROGRAMS and the oracle.

CPU only.  On a box without /root/reference the prebuilt oracle/_ref travels with the snapshot; with neither, skip.
"""
import ctypes as C
import filecmp
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import binding as ob
from oracle import ref_binding as rb
from tests import util

pytestmark = pytest.mark.skipif(not rb.available(), reason="no oracle/_ref build and no reference tree")

ORACLE_DIR = os.path.dirname(os.path.abspath(rb.__file__))
FLAG_SETS = [(), ("tsurf",), ("alb",), ("tsurf", "alb"), ("melt",), ("refr",), ("amp",), ("acc", "smb"), ("shf", "lhf"),
             ("tsurf", "alb", "melt", "refr", "amp", "acc", "smb", "shf", "lhf")]
SCHEMES = ["none", "slater", "isba", "denby", "alex", "unknown"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    same = bits(a) == bits(b)
    both_nan = np.isnan(a) & np.isnan(b)
    bad = ~(same | both_nan)
    assert not bad.any(), "%s: %d of %d values differ, first at %s: %r vs %r" % (
        what, int(bad.sum()), a.size, np.argwhere(bad)[0], a[bad][0], b[bad][0])


def start_state(nx, mask, seed=0):
    S0, m0 = util.init_slab(nx, mask)
    S0[ob.FIDX["amp"]] = 2.5            # external fields the flagged variants keep (bnd%amp, shf, lhf, refr, smb)
    S0[ob.FIDX["shf"]] = 3.0
    S0[ob.FIDX["lhf"]] = -2.0
    S0[ob.FIDX["subl"]] = 1e-9
    S0[ob.FIDX["refr"]] = 1e-9
    S0[ob.FIDX["smb_snow"]] = 2e-9
    return S0, m0


def both(pd, bnd_names, forc, vali, ndays, mask, S0=None):
    nx = forc.shape[2]
    if S0 is None:
        S0, _ = start_state(nx, mask)
    Sa, Sb = S0.copy(), S0.copy()
    m = np.asarray(mask, dtype=np.int32)
    out_a, _ = ob.run(Sa, m.copy(), ob.make_par(pd), ob.make_bnd(bnd_names), forc, ndays, vali=vali, out_days=ndays)
    out_b = rb.run(Sb, m.copy(), pd, bnd_names, forc, ndays, vali=vali, out_days=ndays)
    return (Sa, out_a), (Sb, out_b)


def test_generated_sources_are_regenerated_not_edited(tmp_path):
    """oracle/_ref/*.c must be exactly what the translator emits from the reference tree today"""
    if not rb.have_reference_tree():
        pytest.skip("reference tree not on this box")
    rb.build()
    R = rb.REFERENCE
    jobs = {"semic_ref.c": [], "example_ref.c": [R + "/example/example.f90"],
            "run_particles_ref.c": [R + "/optimize/run_particles.f90"]}
    for name, extra in jobs.items():
        out = tmp_path / name
        subprocess.check_call([sys.executable, os.path.join(ORACLE_DIR, "f90_to_c.py"), "-o", str(out),
                               R + "/surface_physics.f90", R + "/utils.f90"] + extra)
        assert filecmp.cmp(str(out), os.path.join(rb.REF_DIR, name), shallow=False), name
    # and nothing under oracle/_ref is tracked by git
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=util.ROOT, stdout=subprocess.PIPE).stdout.strip()
    assert tracked == b""


def test_translated_output_carries_no_reference_text():
    """the generated C cites file:line, it does not reproduce the Fortran (comments, identifiers aside)"""
    path = os.path.join(rb.REF_DIR, "semic_ref.c")
    if not os.path.isfile(path):
        pytest.skip("generated source did not travel")
    text = open(path).read()
    assert "surface_physics.f90:199" in text and "dmax1" not in text and "intent(" not in text


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("bnd_names", FLAG_SETS, ids=lambda b: "+".join(b) or "noflags")
def test_oracle_equals_reference_transect(scheme, bnd_names):
    forc, vali = util.load_fixture("transect")
    mask = np.array([1, 0, 1, 2, 2, 2, 2], dtype=np.int32)
    pd = dict(util.SCHEME_PARS[scheme], n_ksub=3)
    (Sa, oa), (Sb, ob_) = both(pd, bnd_names, forc, vali, 365, mask)
    assert_bit_equal(oa, ob_, "daily output")
    assert_bit_equal(Sa, Sb, "final state")


@pytest.mark.parametrize("fixture,ndays", [("c01", 365), ("c20", 3652)])
@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("bnd_names", [(), ("tsurf", "alb"), ("melt", "refr", "amp"), ("shf", "lhf", "acc", "smb")],
                         ids=lambda b: "+".join(b) or "noflags")
def test_oracle_equals_reference_stations(fixture, ndays, scheme, bnd_names):
    forc, vali = util.load_fixture(fixture)
    nx = forc.shape[2]
    mask = np.array(([2, 1, 2, 0, 2, 1] * 4)[:nx], dtype=np.int32)
    pd = dict(util.SCHEME_PARS[scheme], n_ksub=3)
    (Sa, oa), (Sb, ob_) = both(pd, bnd_names, forc, vali, ndays, mask)
    assert_bit_equal(oa, ob_, "daily output")
    assert_bit_equal(Sa, Sb, "final state")


@pytest.mark.parametrize("case", ["odd_masks", "amp_zero", "hcrit_zero", "unused_minus_999"])
def test_oracle_equals_reference_corner_cases(case):
    forc, vali = util.load_fixture("transect")
    mask = np.array([1, 0, 1, 2, 2, 2, 2], dtype=np.int32)
    pd = dict(util.PAR_SLATER, n_ksub=3)
    if case == "odd_masks":
        mask = np.array([3, 0, -1, 2, 7, 1, 2], dtype=np.int32)
    elif case == "amp_zero":
        pd["amp"] = 0.0
    elif case == "hcrit_zero":
        pd["hcrit"] = 0.0
    else:
        pd = dict(util.PAR_EXAMPLE, alb_scheme="slater")
    S0, _ = util.init_slab(7, mask)
    (Sa, oa), (Sb, ob_) = both(pd, (), forc, None, 365, mask, S0=S0)
    assert_bit_equal(oa, ob_, "daily output")
    assert_bit_equal(Sa, Sb, "final state")


@pytest.mark.parametrize("scheme,k", [("slater", 3), ("slater", 24), ("isba", 3), ("none", 3), ("alex", 24), ("denby", 1)])
def test_oracle_equals_reference_on_the_synthetic_grid(scheme, k):
    """the SURVEY 8(d) generator (all three masks, all branches of the diurnal cycle) over a full year"""
    from semic_b200 import synth
    nx, nd = 1500, 365
    forc, mask = synth.forcing(nx, nd), synth.mask(nx)
    pd = dict(util.SCHEME_PARS[scheme], n_ksub=k)
    S0, _ = util.init_slab(nx, mask)
    (Sa, oa), (Sb, ob_) = both(pd, (), forc, None, nd, mask, S0=S0)
    assert_bit_equal(oa, ob_, "daily output")
    assert_bit_equal(Sa, Sb, "final state")


def test_elementals_bit_for_bit():
    """the public elementals (surface_physics.f90:378-560) on random in-range and edge arguments"""
    L, R = ob.lib(), rb.lib()
    r = np.random.default_rng(5)
    d = lambda x: C.byref(C.c_double(x))
    for _ in range(2000):
        ts, ta = r.uniform(200., 290.), r.uniform(200., 290.)
        if r.uniform() < 0.1:
            ts = 273.15
        wind, rhoa, qq, sp = r.uniform(0., 25.), r.uniform(0.9, 1.5), r.uniform(0., 8e-3), r.uniform(6e4, 1.05e5)
        o = C.c_double()
        R.ref_sensible_heat_flux(d(ts), d(ta), d(wind), d(rhoa), d(2e-3), d(1000.), C.byref(o))
        assert o.value == L.oracle_sensible_heat_flux(ts, ta, wind, rhoa, 2e-3, 1000.)
        R.ref_longwave_upward(d(ts), d(5.67e-8), C.byref(o))
        assert o.value == L.oracle_longwave_upward(ts, 5.67e-8)
        assert R.ref_ei_sat(d(ts)) == L.oracle_ei_sat(ts) and R.ref_ew_sat(d(ts)) == L.oracle_ew_sat(ts)
        a3, b3 = [C.c_double() for _ in range(3)], [C.c_double() for _ in range(3)]
        R.ref_latent_heat_flux(d(ts), d(wind), d(qq), d(sp), d(rhoa), C.byref(C.c_int(2)), d(5e-4), d(0.62197), d(2.83e6),
                               d(2.5e6), *[C.byref(x) for x in a3])
        L.oracle_latent_heat_flux(ts, wind, qq, sp, rhoa, 2, 5e-4, 0.62197, 2.83e6, 2.5e6, *[C.byref(x) for x in b3])
        assert [x.value for x in a3] == [x.value for x in b3]
        amp, tmean = r.choice([0.0, 1.0, 3.1, 3.5]), r.uniform(-8., 8.)
        if r.uniform() < 0.1:
            tmean = r.choice([amp, -amp, 0.0])
        a2, b2 = [C.c_double(), C.c_double()], [C.c_double(), C.c_double()]
        R.ref_diurnal_cycle(d(amp), d(tmean), C.byref(a2[0]), C.byref(a2[1]))
        L.oracle_diurnal_cycle(amp, tmean, C.byref(b2[0]), C.byref(b2[1]))
        assert_bit_equal([x.value for x in a2], [x.value for x in b2], "diurnal_cycle(%r, %r)" % (amp, tmean))
        R.ref_albedo_slater(C.byref(o), d(ts), d(263.15), d(273.15), d(0.8), d(0.77))
        assert o.value == L.oracle_albedo_slater(ts, 263.15, 273.15, 0.8, 0.77)
        melt, sf = r.choice([0.0, r.uniform(0, 2e-7)]), r.choice([0.0, r.uniform(0, 4e-7)])
        R.ref_albedo_denby(C.byref(o), d(melt), d(0.85), d(0.6), d(6e-8))
        assert o.value == L.oracle_albedo_denby(melt, 0.85, 0.6, 6e-8)
        alb0 = r.uniform(0.55, 0.9)
        x, y = C.c_double(alb0), C.c_double(alb0)
        R.ref_albedo_isba(C.byref(x), d(sf), d(melt), d(86400.), d(86400.), d(0.008), d(0.24), d(15.), d(6e-8), d(0.6), d(0.85))
        L.oracle_albedo_isba(C.byref(y), sf, melt, 86400., 86400., 0.008, 0.24, 15., 6e-8, 0.6, 0.85)
        assert x.value == y.value


def test_crmsd_bit_for_bit():
    """utils.f90:23-42 incl. a constant validation series (sigma = 0 -> Inf/NaN as in the reference)"""
    r = np.random.default_rng(3)
    x1, x2 = r.normal(size=(365, 9)), r.normal(size=(365, 9))
    x2[:, 4] = 1.25
    x1[:, 5] = x2[:, 5] = 0.0
    with np.errstate(all="ignore"):
        assert_bit_equal(ob.calculate_crmsd(x1, x2), rb.calculate_crmsd(x1, x2), "crmsd")
    forc, vali = util.load_fixture("transect")
    for v in range(8):
        assert_bit_equal(ob.calculate_crmsd(forc[:, v, :], vali[:, v, :]), rb.calculate_crmsd(forc[:, v, :], vali[:, v, :]))


def test_average_bit_for_bit():
    """surface_physics_average / field_average (surface_physics.f90:565-629): init, 5 x step, end"""
    r = np.random.default_rng(4)
    nx = 11
    L, R = ob.lib(), rb.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ave_o = np.full((31, nx), 7.0)
    ave_r = rb.state_from_slab(ave_o, np.zeros(nx, dtype=np.int32))
    L.oracle_surface_physics_average(dp(ave_o), dp(ave_o), nx, nx, 0, 0.0)
    R.ref_surface_physics_average(ave_r._p, ave_r._p, b"init", 4, None)
    for _ in range(5):
        now = r.normal(size=(31, nx))
        now_r = rb.state_from_slab(now, np.zeros(nx, dtype=np.int32))
        L.oracle_surface_physics_average(dp(ave_o), dp(now), nx, nx, 1, 0.0)
        R.ref_surface_physics_average(ave_r._p, now_r._p, b"step", 4, None)
        rb.free_state(now_r)
    L.oracle_surface_physics_average(dp(ave_o), dp(ave_o), nx, nx, 2, 5.0)
    R.ref_surface_physics_average(ave_r._p, ave_r._p, b"end", 3, C.byref(C.c_double(5.0)))
    got = np.zeros((31, nx))
    rb.state_to_slab(ave_r, got)
    assert_bit_equal(got, ave_o, "averages")
    rb.free_state(ave_r)


def test_par_load_and_boundary_define_semantics(tmp_path):
    """the reference's own namelist reader semantics, observed on the translated par_load: unspecified keys keep the
    incoming value (Q10), tsticsub = tstic/n_ksub (:810), `d` exponents, space separated strings, unknown boundary names
    ignored (Q11), foreign groups skipped"""
    p = tmp_path / "x.nml"
    p.write_text('&driver\n nloop = 3,\n output = "a/b",\n/\n&surface_physics\n boundary = "tsurf" "massbal" "alb",\n'
                 ' tstic = 8.64d4,\n n_ksub = 24, ! comment\n ceff=2.0e6, alb_scheme = \'slater\'\n/\n')
    par, bnd = rb.Obj("surface_param_class"), rb.Obj("boundary_opt_class")
    par.amp = 3.25
    par.hcrit = 0.5
    b = str(p).encode()
    rb.lib().ref_surface_physics_par_load(par._p, b, len(b))
    assert (par.tstic, par.n_ksub, par.ceff, par.alb_scheme) == (86400.0, 24, 2.0e6, "slater")
    assert par.tsticsub == 86400.0 / 24 and par.amp == 3.25 and par.hcrit == 0.5
    rb.lib().ref_surface_boundary_define(bnd._p, par.addr("boundary"), 30)
    flags = {n: getattr(bnd, n) for n in bnd.components()}
    assert flags == dict(t2m=0, tsurf=1, hsnow=0, alb=1, melt=0, refr=0, smb=0, acc=0, lhf=0, shf=0, subl=0, amp=0)
    # the product's parser agrees on the same file
    from semic_b200 import surface_physics as sp
    mine = sp.Surface_Param_Class()
    mine.amp, mine.hcrit = 3.25, 0.5
    sp.surface_physics_par_load(mine, str(p))
    for n in ("tstic", "tsticsub", "ceff", "amp", "hcrit", "n_ksub"):
        assert getattr(mine, n) == getattr(par, n), n
    assert mine.alb_scheme.strip() == "slater"


def _example_tree(tmp_path, nloop):
    d = tmp_path / "example"
    (d / "data").mkdir(parents=True)
    forc, vali = util.load_fixture("transect")
    # the shipped ASCII tables, rewritten from the committed fixtures (so the test also runs where the reference
    # tree is absent); %.17g round-trips every double exactly
    np.savetxt(d / "data" / "transect_input.txt", forc.reshape(365, -1), fmt="%.17g")
    np.savetxt(d / "data" / "transect_output.txt", vali.reshape(365, -1), fmt="%.17g")
    util.write_example_namelist(str(d / "example.namelist"), nloop=nloop)
    return d


def test_example_program_equals_oracle(tmp_path):
    """`make run_example` with the translated example.f90: namelist -> readers -> init (single-precision 0.8, Q9)
    -> nloop x 365 days -> example.out.  The oracle, driven the same way, must give the same 8 x 7 x 365 numbers."""
    d = _example_tree(tmp_path, nloop=100)
    rb.run_program("example", ["example.namelist"], cwd=str(d))
    got = np.loadtxt(d / "example.out").reshape(365, 8, 7)
    forc, _ = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    S, m = util.init_slab(7, mask, hsnow=5.0, alb=np.float64(np.float32(0.8)))
    out, _ = ob.run(S, m, ob.make_par(util.PAR_EXAMPLE), ob.make_bnd(), forc, 100 * 365, out_days=365)
    assert_bit_equal(got, out, "example.out")
    # ... and it is the committed golden file (tests/golden/make_ref_golden.py)
    gold = np.load(os.path.join(util.GOLDEN, "ref_example_out.npy"))
    assert_bit_equal(got, gold, "golden example.out")


def test_run_particles_program_equals_oracle(tmp_path):
    """optimize/run_particles.f90 translated: three particles in ONE process (state carried from particle to particle,
    boundary flags lag by one particle, Q14), cost = sqrt(sum of squared CRMSDs) written to <prefix>NNNNNN.out"""
    d = _example_tree(tmp_path, nloop=1)
    opt = tmp_path / "optimize"
    opt.mkdir()
    (opt / "optimization.namelist").write_text(
        '&driver\n nloop = 2,\n ntime = 365,\n nx = 7,\n loi_mask(1:3) = 1,1,1,\n'
        ' input_forcing = "../example/data/transect_input.txt",\n validation = "../example/data/transect_output.txt",\n/\n')
    pars = [dict(util.PAR_SLATER, alb_scheme="none", hcrit=0.05 + 0.02 * i, amp=2.0 + 0.7 * i, rcrit=0.3 + 0.2 * i)
            for i in range(3)]
    for i, pd in enumerate(pars):
        rb.write_namelist(str(opt / ("p_%06d.nml" % i)), pd)
    rb.run_program("run_particles", ['"p_"', "0", "2"], cwd=str(opt))
    got = [float(np.loadtxt(opt / ("p_%06d.out" % i))) for i in range(3)]
    # the oracle driven as run_particles.f90:77-82,124-177 does
    forc, vali = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    S, m = ob.new_state(7)
    m[:] = mask
    S[ob.FIDX["hsnow"]] = 1.0
    S[ob.FIDX["alb"]] = vali[0, 1]
    S[ob.FIDX["tsurf"]] = vali[0, 0]
    S[ob.FIDX["alb_snow"]] = vali[0, 1]
    for i, pd in enumerate(pars):
        out, _ = ob.run(S, m, ob.make_par(dict(pd, n_ksub=3)), ob.make_bnd(), forc, 2 * 365, vali=vali, out_days=365)
        assert got[i] == ob.particle_cost(out, vali), i


def test_golden_example_fixture_matches_oracle_on_short_runs():
    """tests/golden/ref_example_nloop1.npy: the translated example program with nloop = 1 (365 days from the cold
    start) -- the golden vector the GPU parity tests use; here it pins the oracle"""
    gold = np.load(os.path.join(util.GOLDEN, "ref_example_nloop1.npy"))
    forc, _ = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    S, m = util.init_slab(7, mask, hsnow=5.0, alb=np.float64(np.float32(0.8)))
    out, _ = ob.run(S, m, ob.make_par(util.PAR_EXAMPLE), ob.make_bnd(), forc, 365, out_days=365)
    assert_bit_equal(out, gold, "golden nloop=1")
