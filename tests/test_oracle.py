"""CPU tests of the oracle (oracle/semic_oracle.c): known answers, SURVEY tripwire vectors, an independent
Python restatement, invariants, and how sensitive the REFERENCE ALGORITHM itself is to rounding (flip context)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from tests import pyref, util


def test_known_answers_elementals():
    L = ob.lib()
    assert L.oracle_ei_sat(273.15) == 611.2 and L.oracle_ew_sat(273.15) == 611.2
    assert abs(L.oracle_longwave_upward(273.15, 5.67e-8) - 315.63697918226688) < 1e-12
    a, b = C.c_double(), C.c_double()
    L.oracle_diurnal_cycle(3.0, 0.0, C.byref(a), C.byref(b))
    assert abs(a.value - 2 * 3.0 / math.pi) < 1e-15 and abs(b.value + 2 * 3.0 / math.pi) < 1e-15
    L.oracle_diurnal_cycle(3.0, -5.0, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (0.0, -5.0)                                               # tmean+amp<0
    L.oracle_diurnal_cycle(3.0, 4.0, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (4.0, 0.0)                                                # tmean>=amp
    smax, smin, tmin, tmax = 0.8, 0.77, 263.15, 273.15
    assert L.oracle_albedo_slater(tmin, tmin, tmax, smax, smin) == smax
    assert abs(L.oracle_albedo_slater(273.15, tmin, tmax, smax, smin) - smin) < 1e-15
    assert abs(L.oracle_albedo_slater(280.0, tmin, tmax, smax, smin) - smin) < 1e-15
    assert L.oracle_albedo_slater(250.0, tmin, tmax, smax, smin) == smax
    assert L.oracle_albedo_slater(272.0, tmin, 270.0, smax, smin) == smax                  # Q12: tmax < ts <= t0 -> tm = 0
    assert L.oracle_albedo_denby(0.0, 0.85, 0.6, 6e-8) == 0.85
    alb = C.c_double(0.7)
    L.oracle_albedo_isba(C.byref(alb), 0.0, 0.0, 86400., 86400., 0.008, 0.24, 15.0, 6e-8, 0.6, 0.85)
    assert abs(alb.value - (0.7 - 0.008)) < 1e-15                                          # dry decline only
    alb = C.c_double(0.602)
    L.oracle_albedo_isba(C.byref(alb), 0.0, 0.0, 86400., 86400., 0.008, 0.24, 15.0, 6e-8, 0.6, 0.85)
    assert alb.value == 0.6                                                                # clamped at alb_smin


def _example_run(pd, nloop):
    """the oracle driven as example/example.f90:74-79,88-127 drives the module"""
    forc, _ = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)                    # example.namelist:5
    S, m = util.init_slab(7, mask, hsnow=5.0, alb=np.float64(np.float32(0.8)))   # Q9: single-precision 0.8
    out, _ = ob.run(S, m, ob.make_par(pd), ob.make_bnd(), forc, nloop * 365, out_days=365)
    return out


@pytest.mark.parametrize("name,scheme,k,nloop", [("out", None, 3, 100), ("nloop1", None, 3, 1), ("slater", "slater", 3, 2),
                                                 ("isba", "isba", 3, 2), ("denby", "denby", 3, 2), ("alex", "alex", 3, 2),
                                                 ("none", "none", 3, 2), ("slater24", "slater", 24, 2)])
def test_oracle_reproduces_the_reference_golden_vectors(name, scheme, k, nloop):
    """tests/golden/ref_example_*.npy were written by the reference's own example program (translated by
    oracle/f90_to_c.py, run in the build container: tests/golden/make_ref_golden.py).  Bit for bit."""
    gold = np.load(os.path.join(util.GOLDEN, "ref_example_%s.npy" % name))
    pd = util.PAR_EXAMPLE if scheme is None else dict(util.SCHEME_PARS[scheme], n_ksub=k)
    out = _example_run(pd, nloop)
    assert np.array_equal(out.view(np.uint64), gold.view(np.uint64))


@pytest.mark.parametrize("label,mask", [("", [1, 1, 1, 2, 2, 2, 2]), ("_ice", [2] * 7)])
def test_oracle_reproduces_the_reference_particle_costs(label, mask):
    """ref_particles_cost.npy: optimize/run_particles.f90 (translated) on the shipped optimization.namelist, three
    particles in one process: state carried over from particle to particle (Q14), nloop = 20"""
    import json
    gold = np.load(os.path.join(util.GOLDEN, "ref_particles_cost%s.npy" % label))
    pars = json.load(open(os.path.join(util.GOLDEN, "ref_particles_pars.json")))
    forc, vali = util.load_fixture("transect")
    S, m = ob.new_state(7)
    m[:] = mask                                                                # optimization.namelist:5 / all ice
    S[ob.FIDX["hsnow"]] = 1.0                                                  # run_particles.f90:77-82
    S[ob.FIDX["alb"]] = vali[0, 1]
    S[ob.FIDX["tsurf"]] = vali[0, 0]
    S[ob.FIDX["alb_snow"]] = vali[0, 1]
    for i, pd in enumerate(pars):
        out, _ = ob.run(S, m, ob.make_par(dict(pd, n_ksub=3)), ob.make_bnd(), forc, 20 * 365, vali=vali, out_days=365)
        assert ob.particle_cost(out, vali) == gold[i]


def test_tripwire_vectors_from_survey():
    """vectors of an independent numpy restatement made at survey time."""
    forc, _ = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    par, bnd = ob.make_par(util.PAR_EXAMPLE), ob.make_bnd()
    S, m = util.init_slab(7, mask, hsnow=5.0, alb=0.8)
    out, _ = ob.run(S, m, par, bnd, forc, 365, out_days=365)
    assert np.allclose(out[:, 0, :].mean(0), [259.0376009662644, 260.0966940039463, 261.47418397788886, 260.2107754852851,
                                              258.1195952250976, 256.8971275080135, 255.60751562073602], rtol=0, atol=2e-10)
    assert np.allclose(S[ob.FIDX["hsnow"]], [3.592153632559457, 3.584472883492908, 3.797000327040003, 4.59841943126199,
                                             4.96228441535274, 5, 5], rtol=0, atol=1e-11)
    assert np.allclose(S[ob.FIDX["hice"]], [0.091656926873103, 0.095097965770272, 0.116362087034888, 0.141432726281011,
                                            0.149143689006251, 0.200507181535675, 0.242823695322179], rtol=0, atol=1e-11)
    assert np.allclose(out[199, 4, :], [1.563685080806398e-07, 1.526783938373903e-07, 9.289032575709848e-08,
                                        3.539342884895535e-08, 0, 0, 0], rtol=1e-9, atol=0)
    S, m = util.init_slab(7, mask, hsnow=5.0, alb=0.8)
    out, _ = ob.run(S, m, par, bnd, forc, 36500, out_days=365)
    trip = [263.14700780751525, 264.47640866970573, 264.87997270472795, 260.740161399957, 258.08253199255205,
            256.8596580925506, 255.56860129142902]
    assert np.abs(out[:, 0, 3:].mean(0) - trip[3:]).max() < 1e-9                            # ice points
    assert np.abs(out[:, 0, :3].mean(0) - trip[:3]).max() < 0.05                            # land: ghost-snow chaos
    S, m = util.init_slab(7, mask, hsnow=1.0, alb=0.8)
    ob.run(S, m, ob.make_par(dict(util.PAR_SLATER, n_ksub=24)), bnd, forc, 365)
    assert np.allclose(S[ob.FIDX["tsurf"]], [229.48829431924685, 230.9572968581782, 239.0762482948204, 243.01332879976127,
                                             240.8521239275544, 240.72466678736458, 240.1431757007083], rtol=0, atol=2e-10)
    assert np.allclose(S[ob.FIDX["hsnow"]], [0.059932622210764, 0.057522685255331, 0.085688325019513, 0.405016262472301,
                                             0.922055427668082, 1.070329879337655, 1.156530516182002], rtol=0, atol=1e-11)


def test_example_run_stays_in_the_band_of_the_shipped_validation_data():
    """The reference ships no SEMIC output, but it ships what its own workflow validates against: MAR fields next to
    every forcing file (example/Readme.md:51-61), compared by example/plot_example.py:57-60 through the Pearson
    correlation and the centred RMSE.  The oracle's `make run_example` twin (example.namelist, 100 loops) must land
    where the tuned model lands: a weak anchor, but one a wrong restatement of the energy balance does not survive."""
    forc, vali = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)                    # example.namelist:5
    S, m = util.init_slab(7, mask, hsnow=5.0, alb=np.float64(np.float32(0.8)))
    out, _ = ob.run(S, m, ob.make_par(util.PAR_EXAMPLE), ob.make_bnd(), forc, 100 * 365, out_days=365)

    def stats(v, p):
        x1, x2 = out[:, v, p], vali[:, v, p]
        r = np.corrcoef(x1, x2)[0, 1]
        crmse = np.sqrt(np.sum(((x1 - x1.mean()) - (x2 - x2.mean())) ** 2) / len(x1))      # plot_example.py:60
        return r, crmse, (x1 - x2).mean()

    for p in range(7):
        r, crmse, bias = stats(0, p)                                           # surface temperature [K]
        assert r > 0.96 and crmse < 4.0 and abs(bias) < 0.6, (p, r, crmse, bias)
        r, crmse, bias = stats(2, p)                                           # net shortwave [W/m2]
        assert r > 0.85 and crmse < 50.0 and abs(bias) < 20.0, (p, r, crmse, bias)
    for p in range(3, 7):                                                      # ice points: melt season is captured
        r, crmse, bias = stats(4, p)
        assert r > 0.7, (p, r)
        r, crmse, bias = stats(7, p)                                           # latent heat flux [W/m2]
        assert r > 0.7 and crmse < 6.0, (p, r, crmse)


@pytest.mark.parametrize("scheme", ["none", "slater", "isba", "denby", "alex", "unknown"])
@pytest.mark.parametrize("bnd_names", [(), ("tsurf", "alb"), ("melt", "refr", "amp"), ("shf", "lhf", "acc", "smb")])
def test_oracle_equals_independent_python_restatement(scheme, bnd_names):
    forc, vali = util.load_fixture("transect")
    mask = np.array([1, 0, 1, 2, 2, 2, 2], dtype=np.int32)
    pd = dict(util.SCHEME_PARS[scheme], n_ksub=3)
    S0, m0 = util.init_slab(7, mask)
    S0[ob.FIDX["amp"]] = 2.5
    S0[ob.FIDX["shf"]] = 3.0
    S0[ob.FIDX["lhf"]] = -2.0
    S0[ob.FIDX["subl"]] = 1e-9
    S0[ob.FIDX["refr"]] = 1e-9
    S0[ob.FIDX["smb_snow"]] = 2e-9
    ndays = 250
    Sa = S0.copy()
    out_a, _ = ob.run(Sa, m0.copy(), ob.make_par(pd), ob.make_bnd(bnd_names), forc, ndays, vali=vali, out_days=ndays)
    Sb = S0.copy()
    out_b = pyref.run(Sb, m0, ob.FIELDS, pd, {n: True for n in bnd_names}, forc, ndays, vali=vali)
    assert np.array_equal(out_a, out_b)          # same operations, same libm: bit for bit
    assert np.array_equal(Sa[:23], Sb[:23])


@pytest.mark.parametrize("case", ["odd_masks", "amp_zero", "hcrit_zero", "unused_minus_999"])
def test_oracle_equals_python_restatement_on_the_corner_cases(case):
    """Q7 (mask values other than 0/1/2 leave alb unchanged), amp = 0 (diurnal_cycle divides by it, :450), hcrit = 0
    (the f_alb exponent divides by epsilon, :338) and the namelists' -999 "unused" values (Q15) put to use by slater"""
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
        pd = dict(util.PAR_EXAMPLE, alb_scheme="slater")             # alb_smin = tmin = -999 (example.namelist:17,20) in use
    S0, m0 = util.init_slab(7, mask)
    Sa = S0.copy()
    out_a, _ = ob.run(Sa, m0.copy(), ob.make_par(pd), ob.make_bnd(()), forc, 200, out_days=200)
    Sb = S0.copy()
    out_b = pyref.run(Sb, m0, ob.FIELDS, pd, {}, forc, 200)
    assert np.array_equal(out_a, out_b, equal_nan=True)
    assert np.array_equal(Sa[:23], Sb[:23], equal_nan=True)
    if case == "odd_masks":
        for i in (0, 2, 4):                                          # alb is never assigned for these points (:350-358)
            assert (out_a[:, 1, i] == S0[ob.FIDX["alb"], i]).all()


def test_invariants_on_random_forcing():
    nx, nd = 500, 400
    forc = util.random_forcing(nx, nd, seed=2)
    r = np.random.default_rng(1)
    mask = r.choice(np.array([0, 1, 2], dtype=np.int32), nx).astype(np.int32)
    S, m = util.init_slab(nx, mask)
    out, br = ob.run(S, m, ob.make_par(util.PAR_SLATER), ob.make_bnd(), forc, nd, out_days=nd, want_branch=True)
    hs = S[ob.FIDX["hsnow"]]
    assert np.isfinite(S).all() and hs.min() >= 0 and hs.max() <= 5.0
    ice, oc = mask == 2, mask == 0
    assert np.array_equal(S[ob.FIDX["melt"]][ice], (S[ob.FIDX["melted_snow"]] + S[ob.FIDX["melted_ice"]])[ice])
    assert (hs[oc] == 0).all() and (S[ob.FIDX["alb"]][oc] == 0.06).all() and (S[ob.FIDX["qmr_res"]][oc] == 0).all()
    assert (out[:, 4, :][:, oc] == 0).all()
    assert (out[:, 0, :][:, ice] <= 273.15).all()
    assert ((br & 4) != 0)[:, ice].all()                        # gate is always on over ice


def test_crmsd_and_cost_match_numpy():
    r = np.random.default_rng(3)
    x1, x2 = r.normal(size=(365, 9)), r.normal(size=(365, 9))
    c = ob.calculate_crmsd(x1, x2)
    ref = np.sqrt((((x1 - x1.mean(0)) - (x2 - x2.mean(0))) ** 2).mean(0)) / x2.std(0)
    assert np.allclose(c, ref, rtol=1e-13)
    out, vali = r.normal(size=(100, 8, 5)), r.normal(size=(100, 8, 5))
    cost = ob.particle_cost(out, vali)
    ref = 0.0
    for v in (0, 4, 3, 2):
        ref += (ob.calculate_crmsd(out[:, v, :], vali[:, v, :]) ** 2).sum()
    assert abs(cost - math.sqrt(ref)) < 1e-12
    assert abs(ob.particle_cost(out, vali, sumsq=True) - ref) < 1e-12


def test_field_average():
    r = np.random.default_rng(4)
    L = ob.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ave = np.full((31, 10), 7.0)
    ref = np.zeros((31, 10))
    assert L.oracle_surface_physics_average(dp(ave), dp(ave), 10, 10, 0, 0.0) == 0
    avg_fields = ["t2m", "tsurf", "hsnow", "alb", "melt", "refr", "smb", "acc", "lhf", "shf", "lwu",
                  "sf", "rf", "sp", "lwd", "swd", "wind", "rhoa", "qq"]
    for _ in range(4):
        now = r.normal(size=(31, 10))
        L.oracle_surface_physics_average(dp(ave), dp(now), 10, 10, 1, 0.0)
        ref += now
    L.oracle_surface_physics_average(dp(ave), dp(ave), 10, 10, 2, 4.0)
    for n in ob.FIELDS:
        i = ob.FIDX[n]
        if n in avg_fields:
            assert np.allclose(ave[i], ref[i] / 4.0, rtol=1e-15)
        else:
            assert (ave[i] == 7.0).all()            # the other 12 fields are not touched (:574-593)
    assert L.oracle_surface_physics_average(dp(ave), dp(ave), 10, 10, 3, 1.0) == -1


def test_reference_algorithm_is_itself_rounding_sensitive(tmp_path):
    """Context for the flip report: two legitimate CPU builds of the SAME restatement (with and without FMA
    contraction, i.e. gfortran -O3 vs -O3 -march=native) disagree on threshold decisions at rates comparable to
    the GPU-vs-oracle flip rate.  The sensitivity is a property of the algorithm (a one-ulp ghost snow layer keeps
    the surface_physics.f90:201 gate on over land), not of the CUDA path."""
    so = tmp_path / "liboracle_fma.so"
    src = os.path.join(os.path.dirname(ob.__file__), "semic_oracle.c")
    rc = subprocess.call(["gcc", "-O3", "-mfma", "-ffp-contract=fast", "-fPIC", "-shared", "-o", str(so), src, "-lm"])
    if rc != 0:
        pytest.skip("gcc -mfma not available")
    L2 = C.CDLL(str(so))
    nx, nd = 6000, 365
    from semic_b200 import synth
    forc = synth.forcing(nx, nd)
    mask = synth.mask(nx)
    par, bnd = ob.make_par(dict(util.PAR_SLATER, n_ksub=3)), ob.make_bnd()
    S, m = util.init_slab(nx, mask)
    out_a, br_a = ob.run(S, m, par, bnd, forc, nd, out_days=nd, want_branch=True)
    S2, m2 = util.init_slab(nx, mask)
    out_b = np.zeros_like(out_a)
    br_b = np.zeros((nd, nx), dtype=np.uint8)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L2.oracle_run.argtypes = ob.lib().oracle_run.argtypes
    L2.oracle_run.restype = None
    L2.oracle_run(dp(S2), nx, m2.ctypes.data_as(C.POINTER(C.c_int)), nx, C.byref(par), C.byref(bnd), dp(forc), None, nd, nx,
                  nd, dp(out_b), nd, nx, nd, br_b.ctypes.data_as(C.POINTER(C.c_ubyte)))
    nflip, pts, _ = util.flip_report(br_b, br_a)
    taint = util.tainted(br_b, br_a)
    errs = util.out_errors(out_b, out_a, exclude=taint)
    print("\n[FMA build vs plain build of the same C restatement] material flips: %d point-days on %d of %d points, "
          "tainted %.4f%%; untainted field-relative error %.2e; by bit %s"
          % (nflip, int(pts.sum()), nx, 100 * taint.mean(), max(errs.values()), util.flips_by_bit(br_b, br_a)))
    assert max(errs.values()) < 1e-10            # away from flips both builds agree to 1e-10
    assert (mask[pts] == 1).all() or pts.sum() < 0.02 * nx
