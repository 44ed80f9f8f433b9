"""Test infrastructure (uses the CPU oracle): prints a parity + flip report (GPU vs oracle) for a few configurations.  Run on the GPU box:
   python tests/parity_report.py [transect|synthetic] [k] [scheme] [strict]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util, gpu_util

BITS = ["ts<t0", "clamp", "gate", "diurnal_lo", "diurnal_hi", "melt>0", "hsnow==0", "snow2ice"]


def report(label, par, forc, mask, ndays, strict, vali=None, bnd=()):
    nx = forc.shape[2]
    S0, m0 = util.init_slab(nx, mask)
    fin_g, out_g, br_g, ms = gpu_util.run_gpu(S0, m0, par, bnd, forc, ndays, vali=vali, out_days=ndays, strict=strict)
    fin_o, out_o, br_o, r_o = gpu_util.run_oracle_diag(S0, m0, par, bnd, forc, ndays, vali=vali, out_days=ndays)
    diff = br_g ^ br_o
    taint = util.tainted(br_g, br_o)
    ill = util.ill_conditioned(r_o, br_o) & ~taint
    oe = util.out_errors(out_g, out_o, exclude=taint)
    oe_all = util.out_errors(out_g, out_o, exclude=taint | ill, pointwise=True)
    oe_ill = util.out_errors(out_g, out_o, exclude=~ill, pointwise=True)
    se = util.state_errors(fin_g, fin_o, exclude_points=taint.any(axis=0))
    print("== %s nx=%d days=%d k=%d scheme=%s strict=%d  kernel %.2f ms" % (label, nx, ndays, par["n_ksub"], par["alb_scheme"], strict, ms))
    print("   daily field-relative (untainted):", " ".join("%s=%.1e" % kv for kv in oe.items()))
    print("   daily pointwise+floor (untainted, well-conditioned):", " ".join("%s=%.1e" % kv for kv in oe_all.items()))
    print("   ill-conditioned point-days (diurnal interior, 1-|tmean/amp| < %.0e): %d of %d (%.4f%%); pointwise+floor on them: %s"
          % (util.ILL_MARGIN, int(ill.sum()), ill.size, 100.0 * ill.mean(), " ".join("%s=%.1e" % kv for kv in oe_ill.items())))
    print("   tainted point-days: %d of %d (%.4f%%) on %d points" % (int(taint.sum()), taint.size, 100.0 * taint.mean(), int(taint.any(axis=0).sum())))
    print("   final state worst: %s" % ", ".join("%s=%.1e" % (k, v) for k, v in sorted(se.items(), key=lambda kv: -kv[1])[:5]))
    print("   flips: %d point-days on %d points; by bit: %s" % (int((diff != 0).sum()), int((diff != 0).any(axis=0).sum()),
          " ".join("%s=%d" % (BITS[b], int(((diff >> b) & 1).sum())) for b in range(8))))
    for v in (3, 4, 5, 6):
        e = util.rel_err(out_g[:, v, :], out_o[:, v, :], util.OUT_FLOOR[v])
        e = np.where(taint | ill, 0.0, e)
        d, i = np.unravel_index(np.argmax(e), e.shape)
        print("   worst %s: day %d pt %d mask %d gpu=%.17g ref=%.17g rel=%.2e | bits=%s | stt=%.17g/%.17g t2m=%.6f swd=%.3f sf=%.3e rf=%.3e prev-day stt diff=%.2e"
              % (util.OUT_NAMES[v], d, i, mask[i], out_g[d, v, i], out_o[d, v, i], e[d, i], bin(br_o[d, i]), out_g[d, 0, i], out_o[d, 0, i],
                 forc[d % forc.shape[0], 8, i], forc[d % forc.shape[0], 2, i], forc[d % forc.shape[0], 0, i], forc[d % forc.shape[0], 1, i],
                 abs(out_g[d-1, 0, i] - out_o[d-1, 0, i]) if d > 0 else 0.0))
    first = (diff != 0)
    if first.any():
        d, i = np.argwhere(first)[0]
        print("   first flip: day %d point %d gpu=%s ref=%s  tsurf gpu=%.17g ref=%.17g" % (d, i, bin(br_g[d, i]), bin(br_o[d, i]), out_g[d, 0, i], out_o[d, 0, i]))
        if d > 0:
            print("      day before: " + " ".join("%s: %.3e" % (util.OUT_NAMES[v], abs(out_g[d-1, v, i] - out_o[d-1, v, i])) for v in range(8)))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "transect"
    ks = [int(sys.argv[2])] if len(sys.argv) > 2 else [3, 24]
    schemes = [sys.argv[3]] if len(sys.argv) > 3 else ["slater", "none", "isba"]
    stricts = [int(sys.argv[4])] if len(sys.argv) > 4 else [0, 1]
    for k in ks:
        for sc in schemes:
            for strict in stricts:
                par = dict(util.SCHEME_PARS[sc], n_ksub=k)
                if what == "transect":
                    forc, _ = util.load_fixture("transect")
                    report("transect", par, forc, np.array([1, 1, 1, 2, 2, 2, 2], dtype=np.int32), 365, strict)
                else:
                    from semic_b200 import engine
                    nx = 20000
                    fs = engine.Series(nx, 9, 365).synth_forcing(20260101)
                    forc = fs.download(); fs.free()
                    r = np.random.default_rng(5)
                    mask = r.choice(np.array([2, 1, 0], dtype=np.int32), size=nx, p=[0.85, 0.10, 0.05]).astype(np.int32)
                    report("synthetic", par, forc, mask, 365, strict)
