"""Parity of the CUDA path (through the C-ABI / the SurfacePhysics module) against the CPU oracle.

Bar (BASELINE.json north_star): FP64 state within 1e-10 relative per variable after one simulated year, with
melt/refreeze threshold flips counted and reported: point-days at or after a point's first branch flip are left
out of the 1e-10 statistic and bounded separately.
"""
import numpy as np
import pytest

from tests import util
from tests.util import TOL

pytestmark = pytest.mark.gpu

PW_TOL_FLUX = 1e-9        # pointwise bound of shf / lhf against their 1e-2 W/m2 floor, see _compare


def _gpu():
    from tests import gpu_util
    return gpu_util


def _compare(par_d, bnd_names, forcing, ndays, mask, vali=None, strict=False, hsnow=1.0, tol=TOL, pw_tol=TOL,
             max_flip_frac=2e-4, label=""):
    g = _gpu()
    nx = forcing.shape[2]
    S0, m0 = util.init_slab(nx, mask, hsnow=hsnow)
    fin_g, out_g, br_g, _ = g.run_gpu(S0, m0, par_d, bnd_names, forcing, ndays, vali=vali, out_days=ndays,
                                      strict=strict)
    fin_o, out_o, br_o, r_o = g.run_oracle_diag(S0, m0, par_d, bnd_names, forcing, ndays, vali=vali, out_days=ndays)
    nflip, pts, first = util.flip_report(br_g, br_o)
    taint = util.tainted(br_g, br_o)
    ill = util.ill_conditioned(r_o, br_o) & ~taint
    oe = util.out_errors(out_g, out_o, exclude=taint)
    se = util.state_errors(fin_g, fin_o, exclude_points=pts)
    pwv = util.out_errors(out_g, out_o, exclude=taint | ill, pointwise=True)
    pw_ill = util.out_errors(out_g, out_o, exclude=~ill, pointwise=True)
    worst = max(list(oe.values()) + list(se.values()))
    print("\n[parity %s] nx=%d days=%d k=%d scheme=%s strict=%d: max rel err %.3e (daily %s); pointwise+floor %s; "
          "ill-conditioned point-days (1-|tmean/amp| < %.0e in the diurnal interior): %d of %d, pointwise on them %.1e; "
          "material flips: %d point-days on %d points (%.4f%% of point-days tainted); by bit %s"
          % (label, nx, ndays, par_d["n_ksub"], par_d["alb_scheme"], strict, worst, max(oe, key=oe.get),
             " ".join("%s=%.1e" % kv for kv in pwv.items()), util.ILL_MARGIN, int(ill.sum()), ill.size, max(pw_ill.values()),
             nflip, int(pts.sum()), 100.0 * taint.mean(), util.flips_by_bit(br_g, br_o)))
    assert worst <= tol, (oe, se)                      # THE bar: per variable, relative to the field (includes ill-conditioned days)
    # The pointwise metric |gpu-ref| / max(|ref|, floor), per variable, on every untainted, well-conditioned point-day: 1e-10
    # as well.  shf and lhf are exchange coefficients times temperature / humidity DIFFERENCES: one ulp of tsurf (5.7e-14 K)
    # times csh*cap*rhoa*wind (up to 64 W/m2/K in the forcing range) is 3.6e-12 W/m2 = 3.6e-10 of the 1e-2 W/m2 floor, so for
    # these two the pointwise bound against the floor is 1e-9 (their error relative to the field is bounded by 1e-10 above).
    for n, e in pwv.items():
        assert e <= (PW_TOL_FLUX if n in ("shf", "lhf") else pw_tol), (n, e, pwv)
    assert max(pw_ill.values()) <= 1e-6 and ill.mean() <= 2e-3, (pw_ill, ill.mean())
    # Flips are threshold events of single points (a one-ulp ghost snow layer over land keeps the :201 gate on),
    # never a systematic drift: ice points (mask 2) are immune, only land points near complete snow melt flip.
    assert (mask[pts] == 1).all() or pts.sum() <= max(2, int(max_flip_frac * nx)), (mask[pts], pts.sum())
    assert taint.mean() <= 0.2, "too many tainted point-days: %.3f" % taint.mean()
    return worst


# ------------------------------------------------------------------------------------------------------------
# shipped forcing (reference fixtures), all schemes, both arithmetic modes
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("scheme", ["none", "slater", "isba", "denby", "alex", "unknown"])
@pytest.mark.parametrize("k", [3, 24])
def test_transect_one_year(scheme, k, strict):
    forc, _ = util.load_fixture("transect")
    par = dict(util.SCHEME_PARS[scheme], n_ksub=k)
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)          # example.namelist:5
    _compare(par, (), forc, 365, mask, strict=strict, label="transect")


def test_example_namelist_config1():
    """config 1 (example/example.f90 with example.namelist): 100 loops x 365 days, k=3, scheme none."""
    forc, _ = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    g = _gpu()
    S0, m0 = util.init_slab(7, mask, hsnow=5.0, alb=np.float64(np.float32(0.8)))    # Q9: single-precision literal
    ndays = 36500
    fin_g, out_g, br_g, _ = g.run_gpu(S0, m0, util.PAR_EXAMPLE, (), forc, ndays, out_days=365)
    fin_o, out_o, br_o = g.run_oracle(S0, m0, util.PAR_EXAMPLE, (), forc, ndays, out_days=365)
    taint = util.tainted(br_g, br_o)
    oe = util.out_errors(out_g, out_o, exclude=taint)
    print("\n[config1] 100 loops: ", {k: "%.2e" % v for k, v in oe.items()}, "flips in last year:", int((br_g != br_o).sum()))
    assert max(oe.values()) <= TOL
    # SURVEY 8(c) tripwire: annual-mean tsurf after 100 loops (computed with alb = 0.8 double; Q9 changes the 9th digit)
    trip = np.array([263.14700780751525, 264.47640866970573, 264.87997270472795, 260.740161399957,
                     258.08253199255205, 256.8596580925506, 255.56860129142902])
    d = np.abs(out_g[:, 0, :].mean(axis=0) - trip)
    assert d[3:].max() < 1e-6          # ice points
    assert d[:3].max() < 0.05          # land points: ghost-snow flips over 100 loops move the annual mean by mK


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("name,scheme,k,nloop", [("out", None, 3, 100), ("nloop1", None, 3, 1), ("slater", "slater", 3, 2),
                                                 ("isba", "isba", 3, 2), ("denby", "denby", 3, 2), ("alex", "alex", 3, 2),
                                                 ("none", "none", 3, 2), ("slater24", "slater", 24, 2)])
def test_against_reference_golden_vectors(name, scheme, k, nloop, strict):
    """tests/golden/ref_example_*.npy: example.out as written by the reference's OWN example program (its Fortran translated
    by oracle/f90_to_c.py and run in the build container on the shipped transect data; tests/golden/make_ref_golden.py).
    The CUDA path, driven like example.f90:74-127, against those files."""
    import os
    gold = np.load(os.path.join(util.GOLDEN, "ref_example_%s.npy" % name))
    forc, _ = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    pd = util.PAR_EXAMPLE if scheme is None else dict(util.SCHEME_PARS[scheme], n_ksub=k)
    g = _gpu()
    S0, m0 = util.init_slab(7, mask, hsnow=5.0, alb=np.float64(np.float32(0.8)))
    ndays = 365 * nloop
    fin_g, out_g, br_g, _ = g.run_gpu(S0, m0, pd, (), forc, ndays, out_days=365, strict=strict)
    fin_o, out_o, br_o = g.run_oracle(S0, m0, pd, (), forc, ndays, out_days=365)
    assert np.array_equal(out_o.view(np.uint64), gold.view(np.uint64))        # the oracle IS the golden file, bit for bit
    taint = util.tainted(br_g, br_o)
    if nloop > 2:
        # flips before the last year are invisible to its bitmap: a point whose daily fields differ at the 1e-6 level
        # went through one (land points near complete snow melt); they are reported, not hidden
        far = (util.rel_err(out_g[:, 0, :], gold[:, 0, :], util.FLOOR["tsurf"]) > 1e-9).any(axis=0)
        taint = taint | far[None, :]
        assert (mask[far] == 1).all(), far
    oe = util.out_errors(out_g, gold, exclude=taint)
    pwv = util.out_errors(out_g, gold, exclude=taint, pointwise=True)
    print("\n[golden %s strict=%d] field-relative %s | pointwise %s | tainted %.3f%%"
          % (name, strict, " ".join("%s=%.1e" % kv for kv in oe.items()), " ".join("%s=%.1e" % kv for kv in pwv.items()),
             100 * taint.mean()))
    assert max(oe.values()) <= TOL, oe
    for n, e in pwv.items():
        assert e <= (PW_TOL_FLUX if n in ("shf", "lhf") else TOL), (n, e)
    assert taint.mean() < 0.5


def test_c20_ten_years():
    """c20 station: 3652 days of real forcing, single point (mask 2)."""
    forc, _ = util.load_fixture("c20")
    _compare(dict(util.PAR_SLATER, n_ksub=3), (), forc, forc.shape[0], np.array([2], dtype=np.int32), label="c20")


# ------------------------------------------------------------------------------------------------------------
# synthetic forcing generated on the device (the bench's generator): oracle consumes the same bytes
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [3, 24])
def test_synthetic_one_year(k):
    from semic_b200 import engine
    nx = 4096 + 333                       # several tiles, ragged tail
    fs = engine.Series(nx, 9, 365).synth_forcing(20260101, point0=0, dayofs=0)
    forc = fs.download()
    fs.free()
    r = np.random.default_rng(5)
    mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.85, 0.10, 0.05]).astype(np.int32)
    _compare(dict(util.PAR_SLATER, n_ksub=k), (), forc, 365, mask, label="synthetic")


# ------------------------------------------------------------------------------------------------------------
# the benchmarked kernel (k_run_lean: 8 output rows, no branch bitmap) against the oracle; the flip taint comes
# from the bitmap run of the same inputs
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scheme,k", [("slater", 3), ("slater", 1), ("slater", 2), ("slater", 5), ("slater", 24),
                                      ("none", 3), ("isba", 3), ("denby", 3), ("alex", 3), ("unknown", 3)])
def test_lean_kernel_parity(scheme, k):
    from semic_b200 import engine
    g = _gpu()
    nx, nwin, ndays, out_days = 2048 + 77, 120, 300, 250         # ragged tail, forcing window wraps, spin-up days unwritten
    fs = engine.Series(nx, 9, nwin).synth_forcing(20260101, point0=31, dayofs=100, daystride=3)
    forc = fs.download()
    fs.free()
    r = np.random.default_rng(7)
    mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.85, 0.10, 0.05]).astype(np.int32)
    par = dict(util.SCHEME_PARS[scheme], n_ksub=k)
    S0, m0 = util.init_slab(nx, mask)
    fin_l, out_l, _, _ = g.run_gpu(S0, m0, par, (), forc, ndays, out_days=out_days, branch=False)     # k_run_lean
    # bitmap run over ALL days: a flip during the unwritten spin-up days taints the point as well
    fin_b, out_b, br_b, _ = g.run_gpu(S0, m0, par, (), forc, ndays, out_days=ndays, branch=True)      # k_run + bitmap
    fin_o, out_o, br_o = g.run_oracle(S0, m0, par, (), forc, ndays, out_days=ndays)
    taint = util.tainted(br_b, br_o)[ndays - out_days:]
    out_b, out_o = out_b[ndays - out_days:], out_o[ndays - out_days:]
    pts = taint.any(axis=0)
    oe = util.out_errors(out_l, out_o, exclude=taint)
    se = util.state_errors(fin_l, fin_o, exclude_points=pts)
    worst = max(list(oe.values()) + list(se.values()))
    same = np.array_equal(out_l, out_b) and np.array_equal(fin_l[:23], fin_b[:23])
    print("\n[lean %s k=%d] max rel err vs oracle %.3e; tainted points %d; bit-identical to the bitmap kernel: %s"
          % (scheme, k, worst, int(pts.sum()), same))
    assert worst <= TOL, (oe, se)
    # the fast translation unit is compiled without implicit FMA contraction and writes its fused operations out, so
    # every instantiation (compile-time n_ksub, unroll factor, run-time scheme switch, bitmap) rounds alike
    assert same


def test_lean_kernel_random_shapes_against_the_bitmap_kernel():
    """shell logic of k_run_lean (tile walk, window wrap, first written day, ragged tail, ring stages across tiles):
    random small shapes must reproduce k_run (selected by the 9-row output) bit for bit"""
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    g = _gpu()
    r = np.random.default_rng(2026)
    for case in range(24):
        nx = int(r.choice([1, 5, 31, 32, 33, 64, 100, 777, 1024, 1025, 5000, 24 * 148 * 32 + 7]))
        nwin = int(r.integers(1, 9))
        ndays = int(r.integers(1, 23))
        out_days = int(r.integers(0, ndays + 1))
        day0 = int(r.integers(0, 3 * nwin))
        k = int(r.choice([1, 2, 3, 3, 3, 8]))
        scheme = str(r.choice(["slater", "none", "isba"]))
        forc = util.random_forcing(nx, nwin, seed=100 + case)
        mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.8, 0.15, 0.05]).astype(np.int32)
        S0, m0 = util.init_slab(nx, mask)
        par, bnd = g.make_par(dict(util.SCHEME_PARS[scheme], n_ksub=k)), g.make_bnd()
        res = []
        for nrow in (8, 9):                                             # 8 rows: k_run_lean, 9 rows: k_run + bitmap
            st = g.make_state(S0, m0)
            fs = engine.Series(nx, 9, nwin).upload(forc)
            outs = engine.Series(nx, nrow, max(out_days, 1))
            engine.run(st, par, bnd, fs, ndays, day0=day0, out=outs if out_days else None, out_days=out_days)
            st.download()
            o = outs.download()[:out_days, :8, :] if out_days else None
            res.append((st.as_slab()[:23].copy(), o))
            spm.surface_dealloc(st)
            fs.free(); outs.free()
        tag = (case, nx, nwin, ndays, out_days, day0, k, scheme)
        assert np.array_equal(res[0][0], res[1][0]), tag
        if out_days:
            assert np.array_equal(res[0][1], res[1][1]), tag


@pytest.mark.parametrize("names", [("tsurf",), ("alb",), ("tsurf", "alb"), ("melt",), ("acc", "smb"), ("shf", "lhf"),
                                   ("tsurf", "alb", "melt", "refr", "amp"), ("smb", "tsurf")])
@pytest.mark.parametrize("k", [3, 4])
def test_lean_kernel_with_external_tsurf_and_alb(names, k):
    """BASELINE configs[3]: isba with bnd%tsurf / bnd%alb fed from an external series runs the plain path with two more
    rows per stage; every other flag set runs the GENERIC physics on the same shell (k_run_lean<EXT = 2>); both are
    bit-identical to the general kernel (9-row run) and within 1e-10 of the oracle"""
    from semic_b200 import engine
    g = _gpu()
    nx, nwin, ndays = 1024 + 99, 40, 150
    forc = util.random_forcing(nx, nwin, seed=21)
    r = np.random.default_rng(5)
    mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.85, 0.10, 0.05]).astype(np.int32)
    S0, m0 = util.init_slab(nx, mask)
    S0[util.ob.FIDX["amp"]] = 2.5          # consulted when bnd%amp
    S0[util.ob.FIDX["shf"]] = 3.0          # frozen when bnd%shf
    S0[util.ob.FIDX["lhf"]] = -2.0
    S0[util.ob.FIDX["subl"]] = 1e-9        # divided by rhow every sub-step under bnd%lhf (Q3)
    S0[util.ob.FIDX["refr"]] = 1e-9
    S0[util.ob.FIDX["smb_snow"]] = 2e-9    # frozen when bnd%smb
    par = dict(util.PAR_ISBA, n_ksub=k)
    # external series: a plain run's own daily output over the window (tsurf <= t0 over snow, alb in [0, 1])
    _, ext, _, _ = g.run_gpu(S0, m0, par, (), forc, nwin, out_days=nwin, branch=False)
    ext = np.ascontiguousarray(ext)
    fin_l, out_l, _, _ = g.run_gpu(S0, m0, par, names, forc, ndays, vali=ext, out_days=ndays, branch=False)   # k_run_lean<EXT>
    fin_b, out_b, br_b, _ = g.run_gpu(S0, m0, par, names, forc, ndays, vali=ext, out_days=ndays, branch=True)  # k_run<GENERIC>
    fin_o, out_o, br_o = g.run_oracle(S0, m0, par, names, forc, ndays, vali=ext, out_days=ndays)
    assert np.array_equal(out_l, out_b) and np.array_equal(fin_l[:23], fin_b[:23])
    taint = util.tainted(br_b, br_o)
    oe = util.out_errors(out_l, out_o, exclude=taint)
    se = util.state_errors(fin_l, fin_o, exclude_points=taint.any(axis=0))
    print("\n[lean ext %s k=%d] daily %.2e state %.2e tainted points %d" % ("+".join(names), k, max(oe.values()),
                                                                            max(se.values()), int(taint.any(axis=0).sum())))
    assert max(oe.values()) <= TOL and max(se.values()) <= TOL


@pytest.mark.parametrize("case", ["odd_masks", "amp_zero", "hcrit_zero", "unused_minus_999"])
def test_lean_kernel_corner_cases(case):
    """Q7 (mask values other than 0/1/2), amp = 0, hcrit = 0 and -999 "unused" parameters put to use, on both kernels"""
    g = _gpu()
    nx, nwin, ndays = 333, 40, 160
    forc = util.random_forcing(nx, nwin, seed=41)
    r = np.random.default_rng(9)
    mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.8, 0.15, 0.05]).astype(np.int32)
    par = dict(util.PAR_SLATER, n_ksub=3)
    if case == "odd_masks":
        mask = r.choice(np.array([2, 1, 0, 3, -1, 7], dtype=np.int32), size=nx).astype(np.int32)
    elif case == "amp_zero":
        par["amp"] = 0.0
    elif case == "hcrit_zero":
        par["hcrit"] = 0.0
    else:
        par = dict(util.PAR_EXAMPLE, alb_scheme="slater")            # alb_smin = tmin = -999
    S0, m0 = util.init_slab(nx, mask)
    fin_l, out_l, _, _ = g.run_gpu(S0, m0, par, (), forc, ndays, out_days=ndays, branch=False)
    fin_b, out_b, br_b, _ = g.run_gpu(S0, m0, par, (), forc, ndays, out_days=ndays, branch=True)
    fin_o, out_o, br_o = g.run_oracle(S0, m0, par, (), forc, ndays, out_days=ndays)
    assert np.array_equal(out_l, out_b, equal_nan=True) and np.array_equal(fin_l[:23], fin_b[:23], equal_nan=True)
    assert np.isfinite(out_l).all() and np.isfinite(out_o).all()
    taint = util.tainted(br_b, br_o)
    oe = util.out_errors(out_l, out_o, exclude=taint)
    se = util.state_errors(fin_l, fin_o, exclude_points=taint.any(axis=0))
    print("\n[corner %s] daily %.2e state %.2e tainted points %d" % (case, max(oe.values()), max(se.values()),
                                                                     int(taint.any(axis=0).sum())))
    assert max(oe.values()) <= TOL and max(se.values()) <= TOL
    if case == "odd_masks":
        odd = ~np.isin(mask, (0, 1, 2))
        assert (out_l[:, 1, odd] == S0[util.ob.FIDX["alb"], odd]).all()     # alb is never assigned (:350-358)


def test_lean_kernel_output_ring_and_chunks():
    """daily output into a ring shorter than the run keeps the last days; a run cut into launches continues bit for bit"""
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    g = _gpu()
    nx, nwin, ndays = 1000, 50, 130
    forc = util.random_forcing(nx, nwin, seed=3)
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::9] = 1
    mask[4::31] = 0
    S0, m0 = util.init_slab(nx, mask)
    par_d = dict(util.PAR_SLATER, n_ksub=3)
    fin_a, out_a, _, _ = g.run_gpu(S0, m0, par_d, (), forc, ndays, out_days=ndays, branch=False)
    # ring of 16 days, all 130 days written: slot (d % 16) holds the last day that mapped to it
    par, bnd = g.make_par(par_d), g.make_bnd()
    st = g.make_state(S0, m0)
    fs = engine.Series(nx, 9, nwin).upload(forc)
    ring = engine.Series(nx, 8, 16)
    engine.run(st, par, bnd, fs, ndays, out=ring, out_days=ndays)
    o = ring.download()
    for slot in range(16):
        d = max(dd for dd in range(ndays) if dd % 16 == slot)
        assert np.array_equal(o[slot], out_a[d]), slot
    st.download()
    assert np.array_equal(st.as_slab()[:23], fin_a[:23])
    spm.surface_dealloc(st)
    # three launches (day0 continues the forcing window): same state and same daily fields as one launch
    st = g.make_state(S0, m0)
    outs = engine.Series(nx, 8, ndays)
    d0 = 0
    got = []
    for n in (1, 49, 80):
        engine.run(st, par, bnd, fs, n, day0=d0, out=outs, out_days=n)
        got.append(outs.download(0, n))
        d0 += n
    st.download()
    assert np.array_equal(st.as_slab()[:23], fin_a[:23])
    assert np.array_equal(np.concatenate(got, axis=0), out_a)
    spm.surface_dealloc(st)
    fs.free(); ring.free(); outs.free()


def test_period_means_in_kernel():
    """opts.avg_days: the rows of `out` are the means of the 8 daily fields over consecutive periods, accumulated
    in registers with the protocol of surface_physics_average (surface_physics.f90:565-629): bit-identical to
    adding the daily rows in order and dividing by the number of days."""
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    g = _gpu()
    nx, nwin, ndays, out_days, period = 1500, 60, 200, 170, 30         # 5 full periods + one of 20 days
    forc = util.random_forcing(nx, nwin, seed=8)
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::6] = 1
    mask[3::40] = 0
    S0, m0 = util.init_slab(nx, mask)
    for k in (3, 4):                                                     # compile-time and run-time n_ksub instantiations
        par_d = dict(util.PAR_SLATER, n_ksub=k)
        fin_a, out_a, _, _ = g.run_gpu(S0, m0, par_d, (), forc, ndays, out_days=out_days, branch=False)
        par, bnd = g.make_par(par_d), g.make_bnd()
        st = g.make_state(S0, m0)
        fs = engine.Series(nx, 9, nwin).upload(forc)
        nper = (out_days + period - 1) // period
        means = engine.Series(nx, 8, nper)
        engine.run(st, par, bnd, fs, ndays, out=means, out_days=out_days, avg_days=period)
        got = means.download()
        st.download()
        assert np.array_equal(st.as_slab()[:23], fin_a[:23])              # the state does not depend on the output mode
        for p in range(nper):
            days = out_a[p * period:(p + 1) * period]
            acc = np.zeros_like(days[0])
            for d in range(days.shape[0]):
                acc = acc + days[d]                                       # field_average 'step' (:614), in day order
            assert np.array_equal(got[p], acc / float(days.shape[0])), (k, p)     # 'end' (:621)
        # overrides / bitmap / strict runs keep daily rows only
        with pytest.raises(Exception):
            engine.run(st, par, g.make_bnd(("tsurf",)), fs, ndays, out=means, out_days=out_days, avg_days=period)
        spm.surface_dealloc(st)
        fs.free(); means.free()


def test_synthetic_ten_years_window_forcing():
    """the bench's C3 protocol at oracle size: 10 years as 10 x 365 days over the 73-day periodic window (one row per
    5 calendar days), slater, n_ksub=3; the 1e-10 bar still holds for the last simulated year"""
    from semic_b200 import engine
    nx, nwin = 768, 73
    fs = engine.Series(nx, 9, nwin).synth_forcing(20260101, point0=4242, dayofs=0, daystride=5)
    forc = fs.download()
    fs.free()
    r = np.random.default_rng(11)
    mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.97, 0.02, 0.01]).astype(np.int32)   # Antarctica-like
    g = _gpu()
    S0, m0 = util.init_slab(nx, mask)
    par = dict(util.PAR_SLATER, n_ksub=3)
    ndays = 3650
    fin_l, out_l, _, _ = g.run_gpu(S0, m0, par, (), forc, ndays, out_days=365, branch=False)          # k_run_lean
    fin_b, out_b, br_b, _ = g.run_gpu(S0, m0, par, (), forc, ndays, out_days=ndays, branch=True)      # bitmap over all days
    fin_o, out_o, br_o = g.run_oracle(S0, m0, par, (), forc, ndays, out_days=ndays)
    taint = util.tainted(br_b, br_o)[ndays - 365:]
    pts = taint.any(axis=0)
    oe = util.out_errors(out_l, out_o[ndays - 365:], exclude=taint)
    se = util.state_errors(fin_l, fin_o, exclude_points=pts)
    print("\n[10 years] last-year max rel err %.3e, final state %.3e, tainted points %d of %d"
          % (max(oe.values()), max(se.values()), int(pts.sum()), nx))
    assert max(oe.values()) <= TOL and max(se.values()) <= TOL
    assert pts.sum() <= 0.05 * nx
    assert np.array_equal(out_l, out_b[ndays - 365:])


def test_synthetic_statistics():
    """the device generator stays inside the ranges of SURVEY 8(d) / the shipped forcing"""
    from semic_b200 import engine
    fs = engine.Series(3000, 9, 365).synth_forcing(20260101, point0=12345, dayofs=0)
    f = fs.download()
    fs.free()
    sf, rf, swd, lwd, wind, sp, rhoa, qq, t2m = [f[:, v, :] for v in range(9)]
    assert np.isfinite(f).all()
    assert sf.min() >= 0 and rf.min() >= 0 and 0.2 < (sf == 0).mean() < 0.95
    assert swd.min() == 0 and 300 < swd.max() <= 400 and 80 < swd.mean() < 200
    assert lwd.min() >= 70 and lwd.max() <= 340
    assert wind.min() >= 0.1 and wind.max() <= 22
    assert 55000 < sp.min() and sp.max() < 108000
    assert 0.7 < rhoa.min() and rhoa.max() < 1.8
    assert 0 < qq.min() and qq.max() < 2e-2
    assert 200 < t2m.min() and t2m.max() < 300
    # counter based: a shifted window reproduces the same values
    fs2 = engine.Series(1000, 9, 100).synth_forcing(20260101, point0=12345 + 500, dayofs=50)
    f2 = fs2.download()
    fs2.free()
    assert np.array_equal(f2, f[50:150, :, 500:1500])


# ------------------------------------------------------------------------------------------------------------
# boundary overrides (generic kernel): example.f90:103-107 feeds tsurf/alb/melt/acc/smb from the validation file
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("names", [("tsurf",), ("alb",), ("tsurf", "alb"), ("melt",), ("acc", "smb"), ("shf", "lhf"),
                                   ("refr",), ("amp",), ("t2m", "hsnow", "subl", "massbal")])
def test_boundary_overrides_isba(names, strict):
    forc, vali = util.load_fixture("transect")
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    g = _gpu()
    S0, m0 = util.init_slab(7, mask)
    S0[util.ob.FIDX["amp"]] = 2.5          # consulted when bnd%amp
    S0[util.ob.FIDX["shf"]] = 3.0          # frozen when bnd%shf
    S0[util.ob.FIDX["lhf"]] = -2.0
    S0[util.ob.FIDX["subl"]] = 1e-9        # divided by rhow every sub-step under bnd%lhf (Q3)
    S0[util.ob.FIDX["refr"]] = 1e-9
    S0[util.ob.FIDX["smb_snow"]] = 2e-9    # frozen when bnd%smb
    par = dict(util.PAR_ISBA, n_ksub=3)
    fin_g, out_g, br_g, _ = g.run_gpu(S0, m0, par, names, forc, 365, vali=vali, out_days=365, strict=strict)
    fin_o, out_o, br_o = g.run_oracle(S0, m0, par, names, forc, 365, vali=vali, out_days=365)
    taint = util.tainted(br_g, br_o)
    oe = util.out_errors(out_g, out_o, exclude=taint)
    se = util.state_errors(fin_g, fin_o, exclude_points=taint.any(axis=0))
    print("\n[bnd %s strict=%d] daily %.2e state %.2e flips %d" % (",".join(names), strict, max(oe.values()),
                                                                    max(se.values()), int((br_g != br_o).sum())))
    assert max(oe.values()) <= TOL and max(se.values()) <= TOL


# ------------------------------------------------------------------------------------------------------------
# the drop-in daily call through the SurfacePhysics module (f2py/test.py call sequence)
# ------------------------------------------------------------------------------------------------------------
def test_dropin_daily_step_module_api(tmp_path):
    import SurfacePhysics as sp
    from oracle import binding as ob
    forc, vali = util.load_fixture("transect")
    nx, ntime = 7, 60
    nml = tmp_path / "test.nml"
    util.write_slater_namelist(str(nml))
    par = sp.surface_physics.Surface_Param_Class()
    sp.surface_physics.surface_physics_par_load(par, str(nml))
    par.nx = nx
    bnd = sp.surface_physics.Boundary_Opt_Class()
    sp.surface_physics.surface_boundary_define(bnd, par.boundary)
    state = sp.surface_physics.Surface_State_Class()
    sp.surface_physics.surface_alloc(state, par.nx)
    state.tsurf = np.array([vali[0, 0, i] for i in range(nx)])
    state.hsnow = np.array([1.0 for i in range(nx)])
    state.hice = np.array([0.0 for i in range(nx)])
    state.alb = vali[0, 1, :]
    state.alb_snow = vali[0, 1, :]
    state.mask = np.array([2 for i in range(nx)])
    state.mask[0:3] = 1
    # oracle twin
    S, m = ob.new_state(nx)
    for n in ("tsurf", "hsnow", "hice", "alb", "alb_snow"):
        S[ob.FIDX[n]] = getattr(state, n)
    m[:] = state.mask
    opar = ob.make_par(dict(util.PAR_SLATER, rcrit=par.rcrit, n_ksub=par.n_ksub))
    obnd = ob.make_bnd()
    names9 = ["sf", "rf", "swd", "lwd", "wind", "sp", "rhoa", "qq", "t2m"]
    worst = 0.0
    for t in range(ntime):
        for v, n in enumerate(names9):
            setattr(state, n, forc[t, v, :])
            S[ob.FIDX[n]] = forc[t, v, :]
        sp.surface_physics.surface_energy_and_mass_balance(state, par, bnd, t, 0)
        ob.step(S, m, opar, obnd)
        e = util.state_errors(state.as_slab(), S)
        worst = max(worst, max(e.values()))
    print("\n[drop-in] 60 daily calls, all 23 fields every day: max rel err %.3e" % worst)
    assert worst <= TOL
    sp.surface_physics.surface_dealloc(state)


def test_dropin_daily_step_pipelined_over_point_slices(monkeypatch):
    """large grids: the per-day call overlaps upload, kernel and download over slices of the grid; same bits as the
    unsliced call, also with a narrowed sync set and for energy_balance / mass_balance alone"""
    from semic_b200 import surface_physics as spm
    from semic_b200._capi import FIELD_ID
    g = _gpu()
    nx, ndays = 5000, 12
    forc = util.random_forcing(nx, ndays, seed=17)
    r = np.random.default_rng(3)
    mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.8, 0.15, 0.05]).astype(np.int32)
    S0, m0 = util.init_slab(nx, mask)
    par, bnd = g.make_par(dict(util.PAR_SLATER, n_ksub=3)), g.make_bnd()
    names9 = ["sf", "rf", "swd", "lwd", "wind", "sp", "rhoa", "qq", "t2m"]
    bit = lambda names: sum(1 << FIELD_ID[x] for x in names)

    def run(slice_pts, narrow):
        if slice_pts:
            monkeypatch.setenv("SEMIC_STEP_SLICE", str(slice_pts))
        else:
            monkeypatch.setenv("SEMIC_STEP_SLICE", str(1 << 30))
        st = g.make_state(S0, m0)
        if narrow:
            # t2m comes back too: energy_balance mutates it (Q1) and the separate mass_balance call uploads it again
            st.set_sync(bit(names9), bit(("tsurf", "alb", "smb", "melt", "acc", "shf", "lhf", "t2m")))
        daily = []
        for t in range(ndays):
            for v, n in enumerate(names9):
                setattr(st, n, forc[t, v, :])
            if t % 3 == 2:                                                   # the two halves called separately
                spm.energy_balance(st, par, bnd, t, 0)
                spm.mass_balance(st, par, bnd, t, 0)
            else:
                spm.surface_energy_and_mass_balance(st, par, bnd, t, 0)
            daily.append(np.stack([np.array(st.tsurf), np.array(st.alb), np.array(st.smb), np.array(st.melt),
                                   np.array(st.acc), np.array(st.shf), np.array(st.lhf)]))
        st.download()
        fin = st.as_slab()[:23].copy()
        spm.surface_dealloc(st)
        return np.array(daily), fin

    ref_d, ref_f = run(0, False)
    for slice_pts, narrow in ((1024, False), (2048, True), (1024, True)):
        d, f = run(slice_pts, narrow)
        assert np.array_equal(d, ref_d), (slice_pts, narrow)
        assert np.array_equal(f, ref_f), (slice_pts, narrow)


def test_energy_and_mass_balance_separately():
    """energy_balance / mass_balance are public module procedures (surface_physics.f90:122)."""
    from oracle import binding as ob
    import ctypes as C
    g = _gpu()
    from semic_b200 import surface_physics as spm
    forc, _ = util.load_fixture("transect")
    mask = np.array([1, 1, 0, 2, 2, 2, 2], dtype=np.int32)
    S0, m0 = util.init_slab(7, mask, tsurf=272.9)
    names9 = ["sf", "rf", "swd", "lwd", "wind", "sp", "rhoa", "qq", "t2m"]
    for v, n in enumerate(names9):
        S0[ob.FIDX[n]] = forc[180, v, :]
    par_d = dict(util.PAR_SLATER, n_ksub=3)
    st = g.make_state(S0, m0)
    par, bnd = g.make_par(par_d), g.make_bnd()
    S = S0.copy()
    now = ob.OraclePar  # noqa: F841
    opar, obnd = ob.make_par(par_d), ob.make_bnd()
    import oracle.binding as b
    L = b.lib()

    OS = b.OracleState
    os_ = OS(7, 7, S.ctypes.data_as(C.POINTER(C.c_double)), m0.ctypes.data_as(C.POINTER(C.c_int)), None, None)
    L.oracle_energy_balance.argtypes = [C.POINTER(OS), C.POINTER(b.OraclePar), C.POINTER(b.OracleBnd)]
    L.oracle_mass_balance.argtypes = [C.POINTER(OS), C.POINTER(b.OraclePar), C.POINTER(b.OracleBnd)]
    for it in range(4):
        spm.energy_balance(st, par, bnd, 1, 0)
        L.oracle_energy_balance(C.byref(os_), C.byref(opar), C.byref(obnd))
        e = util.state_errors(st.as_slab(), S)
        assert max(e.values()) <= TOL, ("energy_balance", it, e)
    spm.mass_balance(st, par, bnd, 1, 0)
    L.oracle_mass_balance(C.byref(os_), C.byref(opar), C.byref(obnd))
    e = util.state_errors(st.as_slab(), S)
    assert max(e.values()) <= TOL, ("mass_balance", e)
    spm.surface_dealloc(st)


# ------------------------------------------------------------------------------------------------------------
# elementals: known-answer tests evaluated by the device functions
# ------------------------------------------------------------------------------------------------------------
def test_elementals_known_answers():
    from semic_b200 import surface_physics as spm
    from oracle import binding as ob
    L = ob.lib()
    assert abs(spm.longwave_upward(273.15, spm.sigm)[0] - 315.63697918226688) < 1e-10
    ts = np.linspace(215., 290., 301)
    lw = spm.longwave_upward(ts, spm.sigm)
    assert np.abs(lw / np.array([L.oracle_longwave_upward(t, spm.sigm) for t in ts]) - 1).max() < 1e-14
    shf = spm.sensible_heat_flux(ts, ts + 1.5, 6.0, 1.2, 2e-3, spm.cap)
    ref = np.array([L.oracle_sensible_heat_flux(t, t + 1.5, 6.0, 1.2, 2e-3, spm.cap) for t in ts])
    assert np.abs(shf - ref).max() <= 1e-12 * np.abs(ref).max()
    lhf, subl, evap = spm.latent_heat_flux(ts, 6.0, 1e-3, 90000., 1.2, 2, 5e-4, spm.eps, spm.cls, spm.clv)
    import ctypes as C
    for i, t in enumerate(ts):
        a, b_, c = C.c_double(), C.c_double(), C.c_double()
        L.oracle_latent_heat_flux(t, 6.0, 1e-3, 90000., 1.2, 2, 5e-4, spm.eps, spm.cls, spm.clv, C.byref(a), C.byref(b_), C.byref(c))
        assert abs(lhf[i] - a.value) <= 1e-12 * max(abs(a.value), 1.0)
        assert abs(subl[i] - b_.value) <= 1e-12 * max(abs(b_.value), 1e-6)
        assert abs(evap[i] - c.value) <= 1e-12 * max(abs(c.value), 1e-6)
        assert (subl[i] == 0.0) == (t >= 273.15) or b_.value == 0.0
    # at t0 both saturation formulas give 611.2 Pa: q = clh*wind*rho*(611.2*eps/(611.2*(eps-1)+sp) - shum)
    lhf0, subl0, evap0 = spm.latent_heat_flux(273.15, 6.0, 0.0, 90000., 1.2, 2, 5e-4, spm.eps, spm.cls, spm.clv)
    q = 5e-4 * 6.0 * 1.2 * (611.2 * spm.eps / (611.2 * (spm.eps - 1.0) + 90000.))
    assert subl0[0] == 0.0 and abs(evap0[0] - q) < 1e-15 and abs(lhf0[0] - q * spm.clv) < 1e-9
    # the one-value forms (what the Fortran shim's ELEMENTAL specifics forward to) return the array forms' bits
    from semic_b200 import _capi
    K = _capi.lib()
    for i in (0, 150, 232, 233, 300):
        t = float(ts[i])
        assert K.semic_longwave_upward_1(t, spm.sigm) == lw[i]
        assert K.semic_sensible_heat_flux_1(t, t + 1.5, 6.0, 1.2, 2e-3, spm.cap) == shf[i]
        got = [K.semic_latent_heat_flux_1(t, 6.0, 1e-3, 90000., 1.2, 2, 5e-4, spm.eps, spm.cls, spm.clv, w) for w in (0, 1, 2)]
        assert got == [lhf[i], subl[i], evap[i]]
    assert np.isnan(K.semic_latent_heat_flux_1(260., 6.0, 1e-3, 90000., 1.2, 2, 5e-4, spm.eps, spm.cls, spm.clv, 3))


# ------------------------------------------------------------------------------------------------------------
# edge cases: empty state, single point, ocean points, zero sub-steps
# ------------------------------------------------------------------------------------------------------------
def test_empty_and_single_point():
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    g = _gpu()
    st = spm.Surface_State_Class()
    spm.surface_alloc(st, 0)
    par, bnd = g.make_par(util.PAR_SLATER), g.make_bnd()
    spm.surface_energy_and_mass_balance(st, par, bnd, 1, 0)          # no-op on nx = 0
    fs = engine.Series(0, 9, 3)
    assert engine.run(st, par, bnd, fs, 3) == 0.0
    fs.free()
    spm.surface_dealloc(st)
    forc, _ = util.load_fixture("c01")
    _compare(dict(util.PAR_SLATER, n_ksub=3), (), forc, 365, np.array([2], dtype=np.int32), label="c01")


def test_ocean_points_invariants():
    forc = util.random_forcing(300, 120, seed=3)
    mask = np.zeros(300, dtype=np.int32)
    g = _gpu()
    S0, m0 = util.init_slab(300, mask)
    fin, out, br, _ = g.run_gpu(S0, m0, util.PAR_SLATER, (), forc, 120, out_days=120)
    from oracle import binding as ob
    assert (fin[ob.FIDX["hsnow"]] == 0).all() and (fin[ob.FIDX["alb"]] == 0.06).all()
    assert (fin[ob.FIDX["qmr_res"]] == 0).all() and (out[:, 4, :] == 0).all()


def test_errors_are_loud(tmp_path):
    from semic_b200 import surface_physics as spm
    from semic_b200 import SemicError
    par = spm.Surface_Param_Class()
    with pytest.raises(SemicError):
        spm.surface_physics_par_load(par, str(tmp_path / "missing.nml"))
    st = spm.Surface_State_Class()
    with pytest.raises(SemicError):
        spm.surface_energy_and_mass_balance(st, par, spm.Boundary_Opt_Class(), 1, 0)     # not allocated
    a, b = spm.Surface_State_Class(), spm.Surface_State_Class()
    spm.surface_alloc(a, 4)
    spm.surface_alloc(b, 4)
    with pytest.raises(SemicError):
        spm.surface_physics_average(a, b, "bogus")
    with pytest.raises(SemicError):
        spm.surface_physics_average(a, b, "end")                                          # nt not provided
    spm.surface_dealloc(a)
    spm.surface_dealloc(b)


def test_surface_physics_average():
    from semic_b200 import surface_physics as spm
    from oracle import binding as ob
    r = np.random.default_rng(9)
    nx = 1500
    now, ave = spm.Surface_State_Class(), spm.Surface_State_Class()
    spm.surface_alloc(now, nx)
    spm.surface_alloc(ave, nx)
    ref = np.zeros((31, nx))
    spm.surface_physics_average(ave, now, "init")
    L = ob.lib()
    import ctypes as C
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for it in range(5):
        slab = r.normal(size=(31, nx))
        for n in ob.FIELDS:
            setattr(now, n, slab[ob.FIDX[n]])
        now.upload()
        spm.surface_physics_average(ave, now, "step")
        L.oracle_surface_physics_average(dp(ref), dp(slab), nx, nx, 1, 0.0)
    spm.surface_physics_average(ave, now, "end", 5.0)
    L.oracle_surface_physics_average(dp(ref), dp(slab), nx, nx, 2, 5.0)
    ave.download()
    assert np.array_equal(ave.as_slab(), ref)            # sums and one division: bit exact
    spm.surface_dealloc(now)
    spm.surface_dealloc(ave)


# ------------------------------------------------------------------------------------------------------------
# host-streamed run == device-resident run, bit for bit
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("slice_pts", [0, 1024])
def test_run_host_equals_run(slice_pts, monkeypatch):
    from semic_b200 import engine
    g = _gpu()
    if slice_pts:
        monkeypatch.setenv("SEMIC_HOST_SLICE", str(slice_pts))       # 3 point slices per chunk of days
    else:
        monkeypatch.delenv("SEMIC_HOST_SLICE", raising=False)
    nx, nwin, ndays = 3000, 30, 75
    forc = util.random_forcing(nx, nwin, seed=11)
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::7] = 1
    S0, m0 = util.init_slab(nx, mask)
    fin_a, out_a, _, _ = g.run_gpu(S0, m0, util.PAR_SLATER, (), forc, ndays, out_days=40, branch=False)
    st = g.make_state(S0, m0)
    par, bnd = g.make_par(util.PAR_SLATER), g.make_bnd()
    hf = engine.PinnedArray((nwin, 9, nx))
    hf.array[...] = forc
    ho = engine.PinnedArray((40, 8, nx))
    ms = engine.run_host(st, par, bnd, hf.array, ndays, h_out=ho.array, out_days=40, chunk_days=8)
    st.download()
    assert ms > 0
    fin_b = st.as_slab()
    assert np.array_equal(fin_a[:23], fin_b[:23])
    assert np.array_equal(out_a, ho.array)
    from semic_b200 import surface_physics as spm
    spm.surface_dealloc(st)
    hf.free()
    ho.free()


def test_run_host_with_period_means(monkeypatch):
    """semic_run_host with opts.avg_days: the streamed run writes the same period means as the device-resident run with
    avg_days, which equal the means of the daily rows formed like surface_physics_average does (sum in day order, / n)"""
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    g = _gpu()
    monkeypatch.setenv("SEMIC_HOST_SLICE", "1024")
    nx, nwin, ndays, out_days, avg, chunk = 2500, 30, 80, 60, 5, 10          # output starts at day 20 = 2 chunks in
    forc = util.random_forcing(nx, nwin, seed=12)
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::5] = 1
    S0, m0 = util.init_slab(nx, mask)
    fin_a, out_a, _, _ = g.run_gpu(S0, m0, util.PAR_SLATER, (), forc, ndays, out_days=out_days, branch=False)
    st = g.make_state(S0, m0)
    par, bnd = g.make_par(util.PAR_SLATER), g.make_bnd()
    hf = engine.PinnedArray((nwin, 9, nx))
    hf.array[...] = forc
    ho = engine.PinnedArray((out_days // avg, 8, nx))
    engine.run_host(st, par, bnd, hf.array, ndays, h_out=ho.array, out_days=out_days, chunk_days=chunk, avg_days=avg)
    st.download()
    assert np.array_equal(fin_a[:23], st.as_slab()[:23])
    want = np.zeros((out_days // avg, 8, nx))
    for p in range(out_days // avg):
        acc = np.zeros((8, nx))
        for d in range(avg):
            acc = acc + out_a[p * avg + d]                                   # field_average 'step' (:614), in day order
        want[p] = acc / float(avg)                                           # 'end' (:621)
    assert np.array_equal(ho.array, want)
    with pytest.raises(Exception):                                           # periods must not straddle pipeline units
        engine.run_host(st, par, bnd, hf.array, ndays, h_out=ho.array, out_days=out_days, chunk_days=8, avg_days=avg)
    spm.surface_dealloc(st)
    hf.free()
    ho.free()


@pytest.mark.parametrize("names", [("tsurf", "alb"), ("melt", "acc")])
def test_run_host_with_external_fields_equals_run(names, monkeypatch):
    """host-streamed run with boundary overrides fed from a HOST validation window (point slices + day chunks)"""
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    g = _gpu()
    monkeypatch.setenv("SEMIC_HOST_SLICE", "1024")
    nx, nwin, ndays = 2500, 12, 31
    forc = util.random_forcing(nx, nwin, seed=51)
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::6] = 1
    S0, m0 = util.init_slab(nx, mask)
    par_d = dict(util.PAR_ISBA, n_ksub=3)
    _, ext, _, _ = g.run_gpu(S0, m0, par_d, (), forc, nwin, out_days=nwin, branch=False)
    ext = np.ascontiguousarray(ext)
    fin_a, out_a, _, _ = g.run_gpu(S0, m0, par_d, names, forc, ndays, vali=ext, out_days=ndays, branch=False)
    st = g.make_state(S0, m0)
    hf = engine.PinnedArray((nwin, 9, nx)); hf.array[...] = forc
    hv = engine.PinnedArray((nwin, 8, nx)); hv.array[...] = ext
    ho = engine.PinnedArray((ndays, 8, nx))
    engine.run_host(st, g.make_par(par_d), g.make_bnd(names), hf.array, ndays, h_vali=hv.array, h_out=ho.array,
                    out_days=ndays, chunk_days=5)
    st.download()
    assert np.array_equal(st.as_slab()[:23], fin_a[:23])
    assert np.array_equal(out_a, ho.array)
    spm.surface_dealloc(st)
    hf.free(); hv.free(); ho.free()


# ------------------------------------------------------------------------------------------------------------
# size-independent properties at large N (BASELINE sizes are too big for the oracle)
# ------------------------------------------------------------------------------------------------------------
def test_large_grid_properties():
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    from oracle import binding as ob
    g = _gpu()
    nx, nwin, ndays = (1 << 24) + 3, 6, 14                    # BASELINE configs[2]: 16.7 M points on one GPU (+ ragged tail)
    par, bnd = g.make_par(dict(util.PAR_SLATER, n_ksub=3)), g.make_bnd()

    def run_shard(p0, n):
        st = spm.Surface_State_Class()
        spm.surface_alloc(st, n)
        from semic_b200._capi import lib, check
        check(lib().semic_state_synth_mask(st.handle, 7, p0, 0.85, 0.10))
        for name, v in (("tsurf", 260.), ("hsnow", 1.0), ("alb", 0.8), ("alb_snow", 0.8)):
            check(lib().semic_state_fill(st.handle, ob.FIDX[name], v))
        fs = engine.Series(n, 9, nwin).synth_forcing(20260101, point0=p0, dayofs=0)
        outs = engine.Series(n, 8, 4)
        engine.run(st, par, bnd, fs, ndays, out=outs, out_days=4)
        st.download()
        fin = st.as_slab()
        mask = np.array(st.mask)
        o = outs.download()
        fs.free(); outs.free(); spm.surface_dealloc(st)
        return fin, o, mask

    fin, out, mask = run_shard(0, nx)
    # invariants of the scheme
    hs = fin[ob.FIDX["hsnow"]]
    assert np.isfinite(fin[:23]).all() and hs.min() >= 0.0 and hs.max() <= 5.0
    ice = mask == 2
    assert np.array_equal(fin[ob.FIDX["melt"]][ice], (fin[ob.FIDX["melted_snow"]] + fin[ob.FIDX["melted_ice"]])[ice])
    oc = mask == 0
    assert (hs[oc] == 0).all() and (fin[ob.FIDX["alb"]][oc] == 0.06).all() and (fin[ob.FIDX["qmr_res"]][oc] == 0).all()
    assert (fin[ob.FIDX["tsurf"]][ice] <= 273.15).all()
    # sharding invariance: any contiguous shard reproduces its slice of the full run bit for bit
    p0, n = 7_777_777, 1_300_001
    fin_s, out_s, _ = run_shard(p0, n)
    assert np.array_equal(fin_s[:23], fin[:23, p0:p0 + n])
    assert np.array_equal(out_s, out[:, :, p0:p0 + n])


# ------------------------------------------------------------------------------------------------------------
# ensemble misfit (config 5): optimize/run_particles.f90 + utils.calculate_crmsd, one all-reduce
# ------------------------------------------------------------------------------------------------------------
def _ensemble_case(nx, ntime, mask):
    from semic_b200 import synth
    forc = synth.forcing(nx, ntime, daystride=3)
    r = np.random.default_rng(0)
    vali = np.ascontiguousarray(r.normal(size=(ntime, 8, nx)) * np.array([5, .1, 30, 1e-8, 1e-8, 1e-8, 5, 5])[None, :, None]
                                + np.array([260, .7, 60, 0, 1e-8, 1e-8, 0, 0])[None, :, None])
    names = ["hcrit", "alb_smax", "amp", "rcrit", "albi", "albl"]                 # optimize/namelist.ranges
    lo = np.array([0.01, 0.80, 0.5, 0.0, 0.35, 0.05])
    hi = np.array([0.5, 0.9, 5.0, 1.0, 0.55, 0.40])
    pop = [{"id": i, "pos": list(lo + (hi - lo) * r.uniform(size=6))} for i in range(13)]
    init = dict(hsnow=1.0, hice=0.0, alb=0.8, tsurf=260.0, alb_snow=0.8)
    return forc, vali, names, pop, init


def test_ensemble_cost_matches_oracle():
    from semic_b200 import driver
    from tests.test_dist_gloo import _oracle_partials
    nx, ntime, nloop = 333, 90, 2
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::9] = 0
    forc, vali, names, pop, init = _ensemble_case(nx, ntime, mask)
    cost_gpu = driver.run_particles(pop, names, forc, vali, mask, init, ntime, nloop)
    cost_ref = driver.run_particles(pop, names, forc, vali, mask, init, ntime, nloop, partial_fn=_oracle_partials)
    err = max(abs(cost_gpu[k] - cost_ref[k]) / cost_ref[k] for k in cost_ref)
    print("\n[ensemble] 13 particles x %d points x %d days x %d loops: max rel cost error %.3e" % (nx, ntime, nloop, err))
    assert err <= 1e-9          # one-pass shifted sums on the device vs the reference's two-pass CRMSD


@pytest.mark.parametrize("label,mask", [("_ice", [2] * 7), ("", [1, 1, 1, 2, 2, 2, 2])])
def test_ensemble_against_the_reference_run_particles_program(label, mask):
    """tests/golden/ref_particles_cost*.npy: the costs the reference's own run_particles program (translated, run in the
    build container) writes on the shipped optimization.namelist (transect, nloop = 20) -- as shipped (three land points)
    and with an all-ice mask.  The first particle starts from the initial state of run_particles.f90:77-82, as every
    particle does here (Q14); the others inherit their predecessor's final state in the reference."""
    import json
    import os
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    g = _gpu()
    gold = np.load(os.path.join(util.GOLDEN, "ref_particles_cost%s.npy" % label))
    pars = json.load(open(os.path.join(util.GOLDEN, "ref_particles_pars.json")))
    forc, vali = util.load_fixture("transect")
    st = spm.Surface_State_Class()
    spm.surface_alloc(st, 7)
    st.mask = np.array(mask, dtype=np.int32)
    st.hsnow = 1.0                                                              # run_particles.f90:77-82
    st.hice = 0.0
    st.alb = vali[0, 1]
    st.tsurf = vali[0, 0]
    st.alb_snow = vali[0, 1]
    st.upload()
    fs = engine.Series(7, 9, 365).upload(forc)
    vs = engine.Series(7, 8, 365).upload(vali)
    plist = [g.make_par(dict(pd, n_ksub=3)) for pd in pars]
    sumsq, _ = engine.ensemble_cost(st, plist, spm.Boundary_Opt_Class(), fs, vs, 365, 20)
    cost = np.sqrt(sumsq)
    fs.free(); vs.free(); spm.surface_dealloc(st)
    rel = np.abs(cost - gold) / gold
    print("\n[ensemble vs reference program%s] cost %r reference %r rel %r" % (label, cost.tolist(), gold.tolist(), rel.tolist()))
    if label == "_ice":
        assert rel[0] <= 1e-9            # one-pass shifted sums vs the reference's two-pass CRMSD
        assert (rel[1:] <= 1e-6).all()   # 20 spin-up loops wash the predecessor's state out
    else:
        # land points near complete snow melt flip between "bare" and a one-ulp ghost snow layer (flips): a
        # cost sums over every point, flipped ones included
        assert (rel <= 1e-3).all()


@pytest.mark.parametrize("case", ["same_scheme_mixed_k", "mixed_scheme_k3", "mixed_both"])
def test_ensemble_with_mixed_schemes_and_substeps(case):
    """particles that differ in alb_scheme and/or n_ksub take the run-time instantiations of the ensemble kernel"""
    from semic_b200 import engine
    from semic_b200 import surface_physics as spm
    from tests.test_dist_gloo import _oracle_partials
    g = _gpu()
    nx, ntime, nloop = 150, 70, 1
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::5] = 1
    forc, vali, names, pop, init = _ensemble_case(nx, ntime, mask)
    schemes = {"same_scheme_mixed_k": ["none"] * 6, "mixed_scheme_k3": ["none", "slater", "isba", "denby", "alex", "slater"],
               "mixed_both": ["slater", "none", "isba", "denby", "alex", "xyz"]}[case]
    ks = {"same_scheme_mixed_k": [3, 2, 5, 3, 4, 3], "mixed_scheme_k3": [3] * 6, "mixed_both": [3, 4, 2, 3, 6, 3]}[case]
    pars = []
    for sch, k in zip(schemes, ks):
        d = dict(util.SCHEME_PARS.get(sch, util.PAR_SLATER), n_ksub=k)
        d["alb_scheme"] = sch
        pars.append(g.make_par(d))
    st = spm.Surface_State_Class()
    spm.surface_alloc(st, nx)
    st.mask = mask
    for k, v in init.items():
        setattr(st, k, v)
    st.upload()
    fs = engine.Series(nx, 9, ntime).upload(forc)
    vs = engine.Series(nx, 8, ntime).upload(vali)
    sumsq, _ = engine.ensemble_cost(st, pars, spm.Boundary_Opt_Class(), fs, vs, ntime, nloop)
    ref = _oracle_partials(pars, forc, vali, mask, init, ntime, nloop)
    fs.free(); vs.free(); spm.surface_dealloc(st)
    err = np.max(np.abs(sumsq - ref) / ref)
    print("\n[ensemble %s] max rel error of the sums of squared CRMSDs %.3e" % (case, err))
    assert err <= 1e-9


def test_ensemble_with_nccl_communicator_single_rank():
    """the library's own NCCL plumbing (unique id -> comm init -> in-place all-reduce) on a 1-rank communicator"""
    import ctypes as C
    from semic_b200 import driver
    from semic_b200._capi import check, lib
    nx, ntime, nloop = 200, 40, 1
    mask = np.full(nx, 2, dtype=np.int32)
    forc, vali, names, pop, init = _ensemble_case(nx, ntime, mask)
    uid = C.create_string_buffer(128)
    check(lib().semic_comm_unique_id(uid))
    comm = C.c_void_p()
    check(lib().semic_comm_init(C.byref(comm), uid, 1, 0))
    a = driver.run_particles(pop, names, forc, vali, mask, init, ntime, nloop, comm=comm)
    b = driver.run_particles(pop, names, forc, vali, mask, init, ntime, nloop)
    check(lib().semic_comm_free(comm))
    assert a == b


# ------------------------------------------------------------------------------------------------------------
# the example program end to end (example/example.f90 with its namelist): files in, example.out out
# ------------------------------------------------------------------------------------------------------------
def test_run_example_program(tmp_path):
    from semic_b200 import driver
    from oracle import binding as ob
    forc, vali = util.load_fixture("transect")
    (tmp_path / "data").mkdir()
    np.savetxt(tmp_path / "data" / "transect_input.txt", forc.reshape(365, -1))
    np.savetxt(tmp_path / "data" / "transect_output.txt", vali.reshape(365, -1))
    util.write_example_namelist(str(tmp_path / "example.namelist"), nloop=3)
    out, now, par, bnd = driver.run_example(str(tmp_path / "example.namelist"))
    # oracle twin of example.f90:74-127
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    S, m = util.init_slab(7, mask, hsnow=5.0, alb=np.float64(np.float32(0.8)))
    ref, br = ob.run(S, m, ob.make_par(util.PAR_EXAMPLE), ob.make_bnd(), forc, 3 * 365, vali=vali, out_days=365)
    written = np.loadtxt(tmp_path / "example.out")                      # plot_example.py reads it the same way
    assert written.shape == (365, 56)
    assert np.array_equal(written.reshape(365, 8, 7), out)              # %24.16E round-trips doubles
    ice = mask == 2
    errs = util.out_errors(out[:, :, ice], ref[:, :, ice])
    assert max(errs.values()) <= TOL, errs
    # the state object holds the last day, like the reference's `surface%now` after the loop
    assert np.allclose(np.array(now.tsurf)[ice], ref[-1, 0, ice], rtol=1e-12)
    assert np.array_equal(np.array(now.sf), forc[364, 0, :]) and np.array_equal(np.array(now.qq), forc[364, 7, :])


# ------------------------------------------------------------------------------------------------------------
# the native command-line drivers: example.x == the Python driver == the oracle; run_particles.x == driver.run_particles
# ------------------------------------------------------------------------------------------------------------
def test_example_x_native_program(tmp_path):
    import os
    import subprocess
    from semic_b200 import driver
    from tests.test_host_api import ROOT, _example_dir
    d = _example_dir(tmp_path, nloop=3)
    r = subprocess.run([os.path.join(ROOT, "semic_b200", "bin", "example.x"), "example.namelist"], cwd=str(d),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "end example program" in r.stdout
    native = np.loadtxt(d / "example.out")                               # plot_example.py reads it the same way
    assert native.shape == (365, 56)
    out, now, par, bnd = driver.run_example(str(d / "example.namelist"), write_output=False)
    assert np.array_equal(native.reshape(365, 8, 7), out)                # same library, same launch: bit for bit


def test_run_particles_x_native_program(tmp_path):
    import os
    import subprocess
    from semic_b200 import driver
    from tests.test_host_api import ROOT, _example_dir
    d = _example_dir(tmp_path, nloop=2)
    forc, vali = util.load_fixture("transect")
    (d / "optimization.namelist").write_text((d / "example.namelist").read_text())
    names = ["hcrit", "alb_smax", "amp", "rcrit", "albi", "albl"]                                # optimize/namelist.ranges
    r = np.random.default_rng(4)
    lo = np.array([0.01, 0.80, 0.5, 0.0, 0.35, 0.05])
    hi = np.array([0.5, 0.9, 5.0, 1.0, 0.55, 0.40])
    pop = [{"id": i, "pos": list(lo + (hi - lo) * r.uniform(size=6))} for i in range(5)]
    for p in pop:                                                        # what optimize/tools.py:7-43 writes
        lines = ["&surface_physics", '  boundary = "", "", "",', "  tstic = 86400.0,", "  ceff = 2.0e6,", "  csh = 2.0e-3,",
                 "  clh = 5.0e-4,", "  tmin = 263.15,", "  tmax = 273.15,", "  afac = -0.18,", "  tmid = 275.35,",
                 "  n_ksub = 3,", '  alb_scheme = "none",']
        lines += ["  %s = %.8f," % (n, v) for n, v in zip(names, p["pos"])]
        (d / ("p_%06d.nml" % p["id"])).write_text("\n".join(lines + ["/", ""]))
    res = subprocess.run([os.path.join(ROOT, "semic_b200", "bin", "run_particles.x"), '"p_"', "0", "9"], cwd=str(d),
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "p_000005.nml' not found, exiting loop" in res.stdout
    native = np.array([float(open(d / ("p_%06d.out" % i)).read()) for i in range(5)])
    mask = np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32)
    init = dict(hsnow=1.0, hice=0.0, alb=vali[0, 1, :], tsurf=vali[0, 0, :], alb_snow=vali[0, 1, :])     # run_particles.f90:77-82
    cost = driver.run_particles(pop, names, forc, vali, mask, init, 365, 2)
    assert np.allclose(native, [cost[i] for i in range(5)], rtol=1e-12, atol=0)


def test_run_from_binary_files_equals_in_memory_run(tmp_path):
    """f3: forcing streamed from a binary columnar file through pinned windows, output streamed to a file"""
    from semic_b200 import io as sio
    from semic_b200 import surface_physics as spm
    g = _gpu()
    nx, nfile, ndays, out_days = 2500, 20, 47, 30
    forc = util.random_forcing(nx, nfile, seed=31)
    mask = np.full(nx, 2, dtype=np.int32)
    mask[::8] = 1
    S0, m0 = util.init_slab(nx, mask)
    par_d = dict(util.PAR_SLATER, n_ksub=3)
    fin_a, out_a, _, _ = g.run_gpu(S0, m0, par_d, (), forc, ndays, out_days=out_days, branch=False)
    sio.write_series(str(tmp_path / "forcing.bin"), forc)
    st = g.make_state(S0, m0)
    ms = sio.run_files(st, g.make_par(par_d), g.make_bnd(), str(tmp_path / "forcing.bin"), ndays,
                       out_path=str(tmp_path / "out.bin"), out_days=out_days, window_days=7, chunk_days=2)
    st.download()
    assert ms > 0
    assert np.array_equal(st.as_slab()[:23], fin_a[:23])
    assert np.array_equal(np.asarray(sio.open_series(str(tmp_path / "out.bin"))), out_a)
    spm.surface_dealloc(st)
