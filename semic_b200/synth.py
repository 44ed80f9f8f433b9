"""Host (numpy) twin of the device forcing generator `semic_series_synth_forcing`.

Same counter-based splitmix64 streams and the same formulas, so the two agree to the last few ulps of
libm's cos/sin/exp/log.  Used where no GPU buffer exists (the CPU reference arm of bench.py, CPU tests).
"""
import numpy as np

V_SF, V_RF, V_SWD, V_LWD, V_WIND, V_SP, V_RHOA, V_QQ, V_T2M = range(9)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uni(seed, point, day, var, j):
    """j-th uniform of the stream keyed by (seed, point, day, var); point/day are broadcastable uint64 arrays."""
    with np.errstate(over="ignore"):
        key = np.uint64(seed) ^ (((point << np.uint64(32)) + day) * np.uint64(16) + np.uint64(var))
        x = _mix64(key + np.uint64(j + 1) * np.uint64(0x9E3779B97F4A7C15))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def gauss4(seed, point, day, var):
    s = uni(seed, point, day, var, 0) + uni(seed, point, day, var, 1) + uni(seed, point, day, var, 2) + \
        uni(seed, point, day, var, 3)
    return (s - 2.0) * 1.7320508075688772


def forcing(npts, ndays, seed=20260101, point0=0, dayofs=0, daystride=1):
    """float64 [ndays][9][npts] in the forcing-file variable order sf rf swd lwd wind sp rhoa qq t2m."""
    pt = (np.arange(npts, dtype=np.uint64) + np.uint64(point0))[None, :]
    dayi = np.arange(ndays, dtype=np.int64) * int(daystride) + int(dayofs)
    day = dayi.astype(np.uint64)[:, None]
    allday = np.uint64(0xFFFFFFFF)
    tbar = 245.0 + 23.0 * uni(seed, pt, allday, 15, 0)
    sp0 = 63000.0 + 37000.0 * uni(seed, pt, allday, 14, 0)
    t0, eps, pi = 273.15, 0.62197, 3.141592653589793
    doy = (dayi % 365)[:, None]
    t2m = tbar + 12.0 * np.cos(2.0 * pi * (doy - 200) / 365.0) + 3.0 * gauss4(seed, pt, day, V_T2M)
    swd = np.where((doy >= 35) & (doy <= 330),
                   np.maximum(0.0, 400.0 * np.sin(pi * (doy - 35) / 295.0)) * (0.6 + 0.4 * uni(seed, pt, day, V_SWD, 0)),
                   0.0)
    lwd = np.minimum(340.0, np.maximum(70.0, 200.0 + 4.0 * (t2m - 260.0) + 15.0 * gauss4(seed, pt, day, V_LWD)))
    wind = np.minimum(22.0, np.maximum(0.1, 6.5 + 3.0 * gauss4(seed, pt, day, V_WIND)))
    sp = sp0 + 1200.0 * gauss4(seed, pt, day, V_SP)
    rhoa = sp / (287.05 * t2m)
    x = t2m - t0
    es = np.where(t2m < t0, 611.2 * np.exp(22.46 * x / (272.62 + x)), 611.2 * np.exp(17.62 * x / (243.12 + x)))
    qsat = es * eps / (es * (eps - 1.0) + sp)
    qq = (0.5 + 0.45 * uni(seed, pt, day, V_QQ, 0)) * qsat
    sf = np.where((t2m < t0) & (uni(seed, pt, day, V_SF, 0) >= 0.65), -3.5e-8 * np.log(1.0 - uni(seed, pt, day, V_SF, 1)), 0.0)
    rf = np.where((t2m > 268.0) & (uni(seed, pt, day, V_RF, 0) >= 0.6), -6.0e-9 * np.log(1.0 - uni(seed, pt, day, V_RF, 1)), 0.0)
    out = np.empty((ndays, 9, npts), dtype=np.float64)
    for v, a in ((V_SF, sf), (V_RF, rf), (V_SWD, swd), (V_LWD, lwd), (V_WIND, wind), (V_SP, sp), (V_RHOA, rhoa),
                 (V_QQ, qq), (V_T2M, t2m)):
        out[:, v, :] = a
    return out


def mask(npts, seed=7, point0=0, frac_ice=0.85, frac_land=0.10):
    pt = np.arange(npts, dtype=np.uint64) + np.uint64(point0)
    u = uni(seed, pt, np.uint64(0xFFFFFFFF), 13, 0)
    return np.where(u < frac_ice, 2, np.where(u < frac_ice + frac_land, 1, 0)).astype(np.int32)
