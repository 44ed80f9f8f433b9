"""Second, independent restatement of SEMIC's time step: scalar Python, one grid point at a time,
written from surface_physics.f90 directly (not from the C oracle).  Slow; used on small cases to pin
the oracle (tests/test_oracle.py).  math.exp/acos/sqrt/pow/tanh are glibc's, like gfortran's."""
import math

import numpy as _np


def _ieee_div(a, b):
    with _np.errstate(all="ignore"):
        return float(_np.float64(a) / _np.float64(b))

T0, SIGM, EPS, CLS, CLM, CLV, CAP, RHOW, HSMAX = 273.15, 5.67e-8, 0.62197, 2.83e6, 3.30e5, 2.5e6, 1000.0, 1000.0, 5.0
EPSIL = 2.220446049250313e-16
PI = 3.141592653589793238462643


def step_point(s, par, bnd, mask):
    """s: dict of the 31 fields of ONE point (modified in place).  surface_physics.f90:138-153"""
    tsticsub = par["tstic"] / float(par["n_ksub"])
    for _ in range(par["n_ksub"]):                                                      # :149-151
        # energy_balance :158-214
        if not bnd.get("shf"):
            s["shf"] = par["csh"] * CAP * s["rhoa"] * s["wind"] * (s["tsurf"] - s["t2m"])   # :387
        if not bnd.get("lhf"):
            s["subl"] = s["evap"] = s["lhf"] = 0.0
            ts = s["tsurf"]
            if ts < T0:
                es = 611.2 * math.exp(22.46 * (ts - T0) / (272.62 + ts - T0))            # :559
                qs = es * EPS / (es * (EPS - 1.0) + s["sp"])
                s["subl"] = par["clh"] * s["wind"] * s["rhoa"] * (qs - s["qq"])
                s["lhf"] = s["subl"] * CLS
            else:
                es = 611.2 * math.exp(17.62 * (ts - T0) / (243.12 + ts - T0))            # :550
                qs = es * EPS / (es * (EPS - 1.0) + s["sp"])
                s["evap"] = par["clh"] * s["wind"] * s["rhoa"] * (qs - s["qq"])
                s["lhf"] = s["evap"] * CLV
        s["subl"] = s["subl"] / RHOW                                                     # :186
        ts = s["tsurf"]
        s["lwu"] = SIGM * ts * ts * ts * ts                                              # :437
        qsb = (1.0 - s["alb"]) * s["swd"] + s["lwd"] - s["lwu"] - s["shf"] - s["lhf"] - s["qmr_res"]   # :194
        s["qmr"] = 0.0
        gate = (mask == 2) or (s["hsnow"] > 0.0)
        if not bnd.get("tsurf"):
            s["tsurf"] = s["tsurf"] + qsb * tsticsub / par["ceff"]                       # :199
            if gate and s["tsurf"] > T0:
                s["qmr"] = (s["tsurf"] - T0) * par["ceff"] / tsticsub
                s["tsurf"] = T0
        if gate:
            s["t2m"] = s["t2m"] + (s["shf"] + s["lhf"]) * tsticsub / par["ceff"]         # :210
    # mass_balance :232-374
    if not bnd.get("amp"):
        s["amp"] = par["amp"]
    amp = s["amp"]
    tmean = s["tsurf"] - T0 + s["qmr"] / par["ceff"] * par["tstic"]
    tmp1 = tmp2 = 0.0
    r = _ieee_div(tmean, amp)                                   # amp = 0: inf or NaN as in Fortran, not an exception
    if abs(r) < 1.0:
        tmp1 = math.acos(r)
        tmp2 = math.sqrt(1.0 - _ieee_div(tmean * tmean, amp * amp))
    if tmean + amp < 0.0:
        below, above = tmean, 0.0
    else:
        above, below = tmean, 0.0
        if abs(tmean) < amp:
            above = (-tmean * tmp1 + amp * tmp2 + PI * tmean) / (PI - tmp1)
            below = (tmean * tmp1 - amp * tmp2) / tmp1
    s["qmr"] = 0.0
    if mask >= 1:
        qmelt = max(0.0, above * par["ceff"] / par["tstic"])
        qcold = max(0.0, abs(below) * par["ceff"] / par["tstic"])
    else:
        qmelt = qcold = 0.0
    if not bnd.get("melt"):
        s["melt"] = qmelt / (RHOW * CLM)
    s["melted_snow"] = min(s["melt"], s["hsnow"] / par["tstic"])
    s["melted_ice"] = s["melt"] - s["melted_snow"]
    if not bnd.get("melt"):
        if mask == 2:
            s["melt"] = s["melted_snow"] + s["melted_ice"]
        else:
            s["melt"] = s["melted_snow"]
            s["melted_ice"] = 0.0
    rr = rs = 0.0
    if not bnd.get("refr"):
        f_rz = par["rcrit"]
        s["refr"] = qcold / (RHOW * CLM)
        rr = min(s["refr"], s["rf"])
        rs = max(s["refr"] - rr, 0.0)
        rs = min(rs, s["melted_snow"])
        rr = f_rz * min(s["hsnow"] / par["tstic"], rr)
        rs = f_rz * min(s["hsnow"] / par["tstic"], rs)
        s["refr"] = rr + rs
        s["qmr"] = s["qmr"] - (1.0 - f_rz) * s["refr"] * RHOW * CLM
    s["runoff"] = s["melt"] + s["rf"] - rr
    if not bnd.get("acc"):
        s["acc"] = s["sf"] - s["subl"] + s["refr"]
    if not bnd.get("smb"):
        s["smb_snow"] = s["sf"] - s["subl"] - s["melted_snow"] + rs
    if mask == 0:
        s["hsnow"] = 0.0
    else:
        s["hsnow"] = max(0.0, s["hsnow"] + s["smb_snow"] * par["tstic"])
    s2i = max(0.0, s["hsnow"] - HSMAX)
    s["hsnow"] = s["hsnow"] - s2i
    s["smb_ice"] = s2i / par["tstic"] - s["melted_ice"] + rr
    s["hice"] = s["hice"] + s["smb_ice"] * par["tstic"]
    if not bnd.get("smb"):
        if mask == 2:
            s["smb"] = s["smb_snow"] + s["smb_ice"] - s2i / par["tstic"]
        else:
            s["smb"] = s["smb_snow"] + max(0.0, s["smb_ice"] - s2i / par["tstic"])
    f_alb = 1.0 - math.exp(-s["hsnow"] / (par["hcrit"] + EPSIL))
    if not bnd.get("alb"):
        sch = par["alb_scheme"].strip()
        if sch == "slater":
            tm = 0.0
            f = 1.0 / (T0 - par["tmin"])
            if par["tmin"] <= s["tsurf"] <= par["tmax"]:
                tm = f * (s["tsurf"] - par["tmin"])
            if s["tsurf"] > T0:
                tm = 1.0
            s["alb_snow"] = par["alb_smax"] - (par["alb_smax"] - par["alb_smin"]) * math.pow(tm, 3.0)
        elif sch == "denby":
            s["alb_snow"] = par["alb_smin"] + (par["alb_smax"] - par["alb_smin"]) * math.exp(-s["melt"] / par["mcrit"])
        elif sch == "isba":
            a = s["alb_snow"]
            tau = par["tstic"]
            dry = a - par["tau_a"] * par["tstic"] / tau
            wet = (a - par["alb_smin"]) * math.exp(-par["tau_f"] * par["tstic"] / tau) + par["alb_smin"]
            new = s["sf"] * par["tstic"] / (par["w_crit"] / RHOW) * (par["alb_smax"] - par["alb_smin"])
            w = 0.0
            if s["melt"] > 0.0:
                w = 1.0 - s["melt"] / par["mcrit"]
            w = min(1.0, max(w, 0.0))
            a = (1.0 - w) * dry + w * wet + new
            s["alb_snow"] = min(par["alb_smax"], max(a, par["alb_smin"]))
        elif sch == "none":
            s["alb_snow"] = par["alb_smax"]
        if mask == 2:
            s["alb"] = par["albi"] + f_alb * (s["alb_snow"] - par["albi"])
        if mask == 1:
            s["alb"] = par["albl"] + f_alb * (s["alb_snow"] - par["albl"])
        if mask == 0:
            s["alb"] = 0.06
        if sch == "alex":
            s["alb_snow"] = par["alb_smin"] + (par["alb_smax"] - par["alb_smin"]) * \
                (0.5 * math.tanh(par["afac"] * (s["t2m"] - par["tmid"])) + 0.5)
            s["alb"] = s["alb_snow"]
    s["qmr_res"] = 0.0 if mask == 0 else s["qmr"]


FORCING_ORDER = ["sf", "rf", "swd", "lwd", "wind", "sp", "rhoa", "qq", "t2m"]
VALI_OVERRIDE = {"tsurf": 0, "alb": 1, "smb": 3, "melt": 4, "acc": 5}


def run(slab, mask, fields, par, bnd, forcing, ndays, vali=None):
    """slab (31,nx) numpy, modified in place; returns out [ndays][8][nx] like example.f90:117-122"""
    import numpy as np
    nx = slab.shape[1]
    nwin = forcing.shape[0]
    out = np.zeros((ndays, 8, nx))
    fid = {n: i for i, n in enumerate(fields)}
    for d in range(ndays):
        for i in range(nx):
            s = {n: float(slab[fid[n], i]) for n in fields}
            for v, n in enumerate(FORCING_ORDER):
                s[n] = float(forcing[d % nwin, v, i])
            if vali is not None:
                for n, v in VALI_OVERRIDE.items():
                    if bnd.get(n):
                        s[n] = float(vali[d % nwin, v, i])
            step_point(s, par, bnd, int(mask[i]))
            for n in fields:
                slab[fid[n], i] = s[n]
            out[d, :, i] = [s["tsurf"], s["alb"], s["swd"] * (1.0 - s["alb"]), s["smb"], s["melt"], s["acc"], s["shf"], s["lhf"]]
    return out
