/*
 * semic_oracle.c -- CPU restatement of SEMIC's daily time step.  TEST INFRASTRUCTURE ONLY:
 * see the header of semic_oracle.h (who may load this, and how it is pinned to the reference).
 *
 * All citations are file:line in /root/reference (mkrapp/semic).  One C loop per Fortran
 * whole-array statement, evaluated left to right exactly as written; automatic arrays are
 * malloc'ed per call like gfortran's.
 */
#include "semic_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* constants, surface_physics.f90:14-25 */
static const double pi    = 3.141592653589793238462643;
static const double t0    = 273.15;
static const double sigm  = 5.67e-8;
static const double eps   = 0.62197;
static const double cls   = 2.83e6;
static const double clm   = 3.30e5;
static const double clv   = 2.5e6;
static const double cap   = 1000.0;
static const double rhow  = 1000.0;
static const double hsmax = 5.0;
static const double epsil = 2.220446049250313e-16; /* epsilon(1.0_dp) */

/* gfortran's dmax1/dmin1 on finite operands; a NaN operand yields the other one */
static inline double dmax1(double a, double b) { return fmax(a, b); }
static inline double dmin1(double a, double b) { return fmin(a, b); }

#define F(now, f) ((now)->slab + (size_t)(f) * (size_t)(now)->ld)

/* ---------------------------------------------------------------- elementals */

/* surface_physics.f90:378-389 */
double oracle_sensible_heat_flux(double ts, double ta, double wind, double rhoa, double csh, double cap_)
{
    return csh * cap_ * rhoa * wind * (ts - ta);                                   /* :387 */
}

/* surface_physics.f90:546-551 */
double oracle_ew_sat(double t)
{
    return 611.2 * exp(17.62 * (t - t0) / (243.12 + t - t0));                      /* :550 */
}

/* surface_physics.f90:555-560 */
double oracle_ei_sat(double t)
{
    return 611.2 * exp(22.46 * (t - t0) / (272.62 + t - t0));                      /* :559 */
}

/* surface_physics.f90:393-430 (the mask argument is unused there too) */
void oracle_latent_heat_flux(double ts, double wind, double shum, double sp, double rhoatm, int mask,
                             double clh, double eps_, double cls_, double clv_,
                             double *lhf, double *subl, double *evap)
{
    double esat_sur, shum_sat;
    (void)mask;
    *subl = 0.0; *evap = 0.0; *lhf = 0.0;                                          /* :410-412 */
    if (ts < t0) {                                                                 /* :415 */
        esat_sur = oracle_ei_sat(ts);
        shum_sat = esat_sur * eps_ / (esat_sur * (eps_ - 1.0) + sp);               /* :418 */
        *subl = clh * wind * rhoatm * (shum_sat - shum);                           /* :420 */
        *lhf  = *subl * cls_;                                                      /* :421 */
    } else {
        esat_sur = oracle_ew_sat(ts);
        shum_sat = esat_sur * eps_ / (esat_sur * (eps_ - 1.0) + sp);               /* :426 */
        *evap = clh * wind * rhoatm * (shum_sat - shum);                           /* :427 */
        *lhf  = *evap * clv_;                                                      /* :428 */
    }
}

/* surface_physics.f90:434-438 */
double oracle_longwave_upward(double ts, double sigma)
{
    return sigma * ts * ts * ts * ts;                                              /* :437 */
}

/* surface_physics.f90:444-472 */
void oracle_diurnal_cycle(double amp, double tmean, double *above, double *below)
{
    double tmp1 = 0.0, tmp2 = 0.0;
    if (fabs(tmean / amp) < 1.0) {                                                 /* :454 */
        tmp1 = acos(tmean / amp);                                                  /* :455 */
        tmp2 = sqrt(1.0 - (tmean * tmean) / (amp * amp));                          /* :456, x**2.0_dp folds to x*x */
    }
    if (tmean + amp < 0.0) {                                                       /* :459 */
        *below = tmean;
        *above = 0.0;
    } else {
        *above = tmean;
        *below = 0.0;
        if (fabs(tmean) < amp) {                                                   /* :465 */
            *above = (-tmean * tmp1 + amp * tmp2 + pi * tmean) / (pi - tmp1);      /* :467 */
            *below = (tmean * tmp1 - amp * tmp2) / tmp1;                           /* :469 */
        }
    }
}

/* surface_physics.f90:476-496 */
void oracle_albedo_isba(double *alb, double sf, double melt, double tstic, double tau, double tau_a,
                        double tau_f, double w_crn, double mcrit, double alb_smin, double alb_smax)
{
    double alb_dry, alb_wet, alb_new, w_alb;
    alb_dry = *alb - tau_a * tstic / tau;                                          /* :484 */
    alb_wet = (*alb - alb_smin) * exp(-tau_f * tstic / tau) + alb_smin;            /* :486 */
    alb_new = sf * tstic / (w_crn / rhow) * (alb_smax - alb_smin);                 /* :487 */
    w_alb = 0.0;
    if (melt > 0.0) w_alb = 1.0 - melt / mcrit;                                    /* :491 */
    w_alb = dmin1(1.0, dmax1(w_alb, 0.0));                                         /* :492 */
    *alb = (1.0 - w_alb) * alb_dry + w_alb * alb_wet + alb_new;                    /* :493 */
    *alb = dmin1(alb_smax, dmax1(*alb, alb_smin));                                 /* :495 */
}

/* surface_physics.f90:503-530 */
double oracle_albedo_slater(double tsurf, double tmin, double tmax, double alb_smax, double alb_smin)
{
    double tm = 0.0, f;
    f = 1.0 / (t0 - tmin);                                                         /* :518 */
    if (tsurf >= tmin && tsurf <= tmax) tm = f * (tsurf - tmin);                   /* :519-521 */
    if (tsurf > t0) tm = 1.0;                                                      /* :522-524 */
    return alb_smax - (alb_smax - alb_smin) * pow(tm, 3.0);                        /* :529, real exponent -> libm pow */
}

/* surface_physics.f90:535-542 */
double oracle_albedo_denby(double melt, double alb_smax, double alb_smin, double mcrit)
{
    return alb_smin + (alb_smax - alb_smin) * exp(-melt / mcrit);                  /* :541 */
}

/* ---------------------------------------------------------------- energy balance */

/* surface_physics.f90:158-214 */
void oracle_energy_balance(oracle_state *now, const oracle_par *par, const oracle_bnd *bnd)
{
    const int nx = now->nx;
    int i;
    double *tsurf = F(now, OF_TSURF), *t2m = F(now, OF_T2M), *wind = F(now, OF_WIND), *rhoa = F(now, OF_RHOA);
    double *shf = F(now, OF_SHF), *lhf = F(now, OF_LHF), *subl = F(now, OF_SUBL), *evap = F(now, OF_EVAP);
    double *qq = F(now, OF_QQ), *sp = F(now, OF_SP), *lwu = F(now, OF_LWU), *alb = F(now, OF_ALB);
    double *swd = F(now, OF_SWD), *lwd = F(now, OF_LWD), *qmr_res = F(now, OF_QMR_RES), *qmr = F(now, OF_QMR);
    double *hsnow = F(now, OF_HSNOW);
    const int *mask = now->mask;
    double *qsb = (double *)malloc(sizeof(double) * (size_t)(nx > 0 ? nx : 1));    /* automatic array :167 */

    if (now->branch && !bnd->lhf) {
        /* flip report: the ts<t0 decision of latent_heat_flux (:415) is taken on tsurf BEFORE :199 */
        for (i = 0; i < nx; ++i) {
            if (tsurf[i] < t0) now->branch[i] |= OB_TS_LT_T0;
            else now->branch[i] &= (unsigned char)~OB_TS_LT_T0;
        }
    }
    if (!bnd->shf) {                                                               /* :171 */
        for (i = 0; i < nx; ++i) shf[i] = 0.0;                                     /* :172 */
        for (i = 0; i < nx; ++i)                                                   /* :173 */
            shf[i] = oracle_sensible_heat_flux(tsurf[i], t2m[i], wind[i], rhoa[i], par->csh, cap);
    }
    if (!bnd->lhf) {                                                               /* :179 */
        for (i = 0; i < nx; ++i) subl[i] = 0.0;                                    /* :180 */
        for (i = 0; i < nx; ++i) evap[i] = 0.0;                                    /* :181 */
        for (i = 0; i < nx; ++i) lhf[i] = 0.0;                                     /* :182 */
        for (i = 0; i < nx; ++i)                                                   /* :183-184 */
            oracle_latent_heat_flux(tsurf[i], wind[i], qq[i], sp[i], rhoa[i], mask[i],
                                    par->clh, eps, cls, clv, &lhf[i], &subl[i], &evap[i]);
    }
    for (i = 0; i < nx; ++i) subl[i] = subl[i] / rhow;                             /* :186 (always) */

    for (i = 0; i < nx; ++i) lwu[i] = 0.0;                                         /* :190 */
    for (i = 0; i < nx; ++i) lwu[i] = oracle_longwave_upward(tsurf[i], sigm);      /* :191 */

    for (i = 0; i < nx; ++i)                                                       /* :194 */
        qsb[i] = (1.0 - alb[i]) * swd[i] + lwd[i] - lwu[i] - shf[i] - lhf[i] - qmr_res[i];

    for (i = 0; i < nx; ++i) qmr[i] = 0.0;                                         /* :197 */
    if (!bnd->tsurf) {                                                             /* :198 */
        for (i = 0; i < nx; ++i) tsurf[i] = tsurf[i] + qsb[i] * par->tsticsub / par->ceff;   /* :199 */
        for (i = 0; i < nx; ++i) {                                                 /* :201-204 */
            if ((mask[i] == 2 || hsnow[i] > 0.0) && tsurf[i] > t0) {
                qmr[i]   = (tsurf[i] - t0) * par->ceff / par->tsticsub;
                tsurf[i] = t0;
            }
        }
    }
    if (now->branch) {
        /* flip report: decisions of this sub-step (the last sub-step's survive) */
        for (i = 0; i < nx; ++i) {
            unsigned char b = (unsigned char)(now->branch[i] & ~(OB_CLAMP | OB_GATE));
            if (mask[i] == 2 || hsnow[i] > 0.0) b |= OB_GATE;
            if (qmr[i] != 0.0) b |= OB_CLAMP;
            now->branch[i] = b;
        }
    }
    for (i = 0; i < nx; ++i) {                                                     /* :209-211 */
        if (mask[i] == 2 || hsnow[i] > 0.0)
            t2m[i] = t2m[i] + (shf[i] + lhf[i]) * par->tsticsub / par->ceff;
    }
    free(qsb);
}

/* ---------------------------------------------------------------- mass balance */

static void rtrim_copy(char *dst, const char *src, size_t n)
{
    size_t k = 0;
    while (k + 1 < n && src[k] != '\0') { dst[k] = src[k]; ++k; }
    dst[k] = '\0';
    while (k > 0 && dst[k - 1] == ' ') dst[--k] = '\0';                            /* trim() */
}

/* surface_physics.f90:232-374 */
void oracle_mass_balance(oracle_state *now, const oracle_par *par, const oracle_bnd *bnd)
{
    const int nx = now->nx;
    const size_t n = (size_t)(nx > 0 ? nx : 1);
    int i;
    double *tsurf = F(now, OF_TSURF), *t2m = F(now, OF_T2M), *amp = F(now, OF_AMP), *qmr = F(now, OF_QMR);
    double *melt = F(now, OF_MELT), *melted_snow = F(now, OF_MELTED_SNOW), *melted_ice = F(now, OF_MELTED_ICE);
    double *hsnow = F(now, OF_HSNOW), *hice = F(now, OF_HICE), *refr = F(now, OF_REFR), *rf = F(now, OF_RF);
    double *sf = F(now, OF_SF), *subl = F(now, OF_SUBL), *runoff = F(now, OF_RUNOFF), *acc = F(now, OF_ACC);
    double *smb_snow = F(now, OF_SMB_SNOW), *smb_ice = F(now, OF_SMB_ICE), *smb = F(now, OF_SMB);
    double *alb = F(now, OF_ALB), *alb_snow = F(now, OF_ALB_SNOW), *qmr_res = F(now, OF_QMR_RES);
    const int *mask = now->mask;
    char scheme[256];
    /* automatic arrays :242-243 */
    double *qmelt = (double *)malloc(sizeof(double) * n), *qcold = (double *)malloc(sizeof(double) * n);
    double *below = (double *)malloc(sizeof(double) * n), *above = (double *)malloc(sizeof(double) * n);
    double *f_rz = (double *)malloc(sizeof(double) * n), *f_alb = (double *)malloc(sizeof(double) * n);
    double *refrozen_rain = (double *)malloc(sizeof(double) * n);
    double *refrozen_snow = (double *)malloc(sizeof(double) * n);
    double *snow_to_ice = (double *)malloc(sizeof(double) * n);

    /* Q5: refrozen_rain/snow are uninitialised in the reference when bnd%refr; defined as 0 here */
    for (i = 0; i < nx; ++i) { refrozen_rain[i] = 0.0; refrozen_snow[i] = 0.0; }

    if (!bnd->amp) {                                                               /* :246-248 */
        for (i = 0; i < nx; ++i) amp[i] = par->amp;
    }
    for (i = 0; i < nx; ++i)                                                       /* :249 */
        oracle_diurnal_cycle(amp[i], tsurf[i] - t0 + qmr[i] / par->ceff * par->tstic, &above[i], &below[i]);
    if (now->rdiag) {
        /* conditioning report: r = tmean/amp of the diurnal cycle (:454).  For |r| -> 1 inside the interior branch the
         * numerators of :467/:469 cancel to O((1-|r|)^1.5): the reference's own result carries rounding noise of relative
         * size ~1e-15/(1-|r|)^1.5 there */
        for (i = 0; i < nx; ++i) now->rdiag[i] = (tsurf[i] - t0 + qmr[i] / par->ceff * par->tstic) / amp[i];
    }
    if (now->branch) {
        for (i = 0; i < nx; ++i) {
            double tmean = tsurf[i] - t0 + qmr[i] / par->ceff * par->tstic;
            unsigned char b = (unsigned char)(now->branch[i] & ~(OB_DIURNAL_LO | OB_DIURNAL_HI));
            if (tmean + amp[i] < 0.0) { /* id 0 */ }
            else if (fabs(tmean) < amp[i]) b |= OB_DIURNAL_LO;     /* id 1 */
            else b |= OB_DIURNAL_HI;                               /* id 2 */
            now->branch[i] = b;
        }
    }
    for (i = 0; i < nx; ++i) qmr[i] = 0.0;                                         /* :250 */

    for (i = 0; i < nx; ++i) {                                                     /* :252-261 */
        if (mask[i] >= 1) {
            qmelt[i] = dmax1(0.0, above[i] * par->ceff / par->tstic);              /* :254 */
            qcold[i] = dmax1(0.0, fabs(below[i]) * par->ceff / par->tstic);        /* :257 */
        } else {
            qmelt[i] = 0.0;
            qcold[i] = 0.0;
        }
    }

    if (!bnd->melt) {                                                              /* :264 */
        for (i = 0; i < nx; ++i) melt[i] = qmelt[i] / (rhow * clm);                /* :266 */
    }
    for (i = 0; i < nx; ++i) melted_snow[i] = dmin1(melt[i], hsnow[i] / par->tstic);   /* :269 */
    for (i = 0; i < nx; ++i) melted_ice[i] = melt[i] - melted_snow[i];             /* :270 */

    if (!bnd->melt) {                                                              /* :272 */
        for (i = 0; i < nx; ++i) {                                                 /* :274-279 */
            if (mask[i] == 2) {
                melt[i] = melted_snow[i] + melted_ice[i];
            } else {
                melt[i] = melted_snow[i];
                melted_ice[i] = 0.0;
            }
        }
    }

    if (!bnd->refr) {                                                              /* :283 */
        for (i = 0; i < nx; ++i) f_rz[i] = par->rcrit;                             /* :284 */
        for (i = 0; i < nx; ++i) refr[i] = qcold[i] / (rhow * clm);                /* :286 */
        for (i = 0; i < nx; ++i) refrozen_rain[i] = dmin1(refr[i], rf[i]);         /* :287 */
        for (i = 0; i < nx; ++i) refrozen_snow[i] = dmax1(refr[i] - refrozen_rain[i], 0.0);   /* :289 */
        for (i = 0; i < nx; ++i) refrozen_snow[i] = dmin1(refrozen_snow[i], melted_snow[i]);  /* :291 */
        for (i = 0; i < nx; ++i) refrozen_rain[i] = f_rz[i] * dmin1(hsnow[i] / par->tstic, refrozen_rain[i]);  /* :293 */
        for (i = 0; i < nx; ++i) refrozen_snow[i] = f_rz[i] * dmin1(hsnow[i] / par->tstic, refrozen_snow[i]);  /* :294 */
        for (i = 0; i < nx; ++i) refr[i] = refrozen_rain[i] + refrozen_snow[i];    /* :295 */
        for (i = 0; i < nx; ++i) qmr[i] = qmr[i] - (1.0 - f_rz[i]) * refr[i] * rhow * clm;   /* :298 */
    }

    for (i = 0; i < nx; ++i) runoff[i] = 0.0;                                      /* :302 */
    for (i = 0; i < nx; ++i) runoff[i] = melt[i] + rf[i] - refrozen_rain[i];       /* :303 */

    if (!bnd->acc) {                                                               /* :306 */
        for (i = 0; i < nx; ++i) acc[i] = sf[i] - subl[i] + refr[i];               /* :307 */
    }
    if (!bnd->smb) {                                                               /* :311 */
        for (i = 0; i < nx; ++i) smb_snow[i] = sf[i] - subl[i] - melted_snow[i] + refrozen_snow[i];   /* :312 */
    }

    for (i = 0; i < nx; ++i) {                                                     /* :315-320 */
        if (mask[i] == 0) hsnow[i] = 0.0;
        else hsnow[i] = dmax1(0.0, hsnow[i] + smb_snow[i] * par->tstic);           /* :319 */
    }

    for (i = 0; i < nx; ++i) snow_to_ice[i] = dmax1(0.0, hsnow[i] - hsmax);        /* :323 */
    for (i = 0; i < nx; ++i) hsnow[i] = hsnow[i] - snow_to_ice[i];                 /* :324 */
    for (i = 0; i < nx; ++i) smb_ice[i] = snow_to_ice[i] / par->tstic - melted_ice[i] + refrozen_rain[i];   /* :325 */
    for (i = 0; i < nx; ++i) hice[i] = hice[i] + smb_ice[i] * par->tstic;          /* :326 */

    if (!bnd->smb) {                                                               /* :329 */
        for (i = 0; i < nx; ++i) {                                                 /* :330-334 */
            if (mask[i] == 2)
                smb[i] = smb_snow[i] + smb_ice[i] - snow_to_ice[i] / par->tstic;
            else
                smb[i] = smb_snow[i] + dmax1(0.0, smb_ice[i] - snow_to_ice[i] / par->tstic);
        }
    }

    for (i = 0; i < nx; ++i) f_alb[i] = 1.0 - exp(-hsnow[i] / (par->hcrit + epsil));   /* :338 */
    rtrim_copy(scheme, par->alb_scheme, sizeof scheme);
    if (!bnd->alb) {                                                               /* :339 */
        if (strcmp(scheme, "slater") == 0) {                                       /* :340-341 */
            for (i = 0; i < nx; ++i)
                alb_snow[i] = oracle_albedo_slater(tsurf[i], par->tmin, par->tmax, par->alb_smax, par->alb_smin);
        } else if (strcmp(scheme, "denby") == 0) {                                 /* :342-343 */
            for (i = 0; i < nx; ++i)
                alb_snow[i] = oracle_albedo_denby(melt[i], par->alb_smax, par->alb_smin, par->mcrit);
        } else if (strcmp(scheme, "isba") == 0) {                                  /* :344-346 */
            for (i = 0; i < nx; ++i)
                oracle_albedo_isba(&alb_snow[i], sf[i], melt[i], par->tstic, par->tstic, par->tau_a, par->tau_f,
                                   par->w_crit, par->mcrit, par->alb_smin, par->alb_smax);
        } else if (strcmp(scheme, "none") == 0) {                                  /* :347-348 */
            for (i = 0; i < nx; ++i) alb_snow[i] = par->alb_smax;
        }
        for (i = 0; i < nx; ++i)                                                   /* :350-352 */
            if (mask[i] == 2) alb[i] = par->albi + f_alb[i] * (alb_snow[i] - par->albi);
        for (i = 0; i < nx; ++i)                                                   /* :353-355 */
            if (mask[i] == 1) alb[i] = par->albl + f_alb[i] * (alb_snow[i] - par->albl);
        for (i = 0; i < nx; ++i)                                                   /* :356-358 */
            if (mask[i] == 0) alb[i] = 0.06;
        if (strcmp(scheme, "alex") == 0) {                                         /* :360-363 */
            for (i = 0; i < nx; ++i)
                alb_snow[i] = par->alb_smin + (par->alb_smax - par->alb_smin)
                              * (0.5 * tanh(par->afac * (t2m[i] - par->tmid)) + 0.5);
            for (i = 0; i < nx; ++i) alb[i] = alb_snow[i];
        }
    }

    for (i = 0; i < nx; ++i) {                                                     /* :368-372 */
        if (mask[i] == 0) qmr_res[i] = 0.0;
        else qmr_res[i] = qmr[i];
    }

    if (now->branch) {
        for (i = 0; i < nx; ++i) {
            unsigned char b = (unsigned char)(now->branch[i] & ~(OB_MELT_POS | OB_HSNOW_ZERO | OB_SNOW2ICE));
            if (melt[i] > 0.0) b |= OB_MELT_POS;
            if (hsnow[i] == 0.0) b |= OB_HSNOW_ZERO;
            if (snow_to_ice[i] > 0.0) b |= OB_SNOW2ICE;
            now->branch[i] = b;
        }
    }

    free(qmelt); free(qcold); free(below); free(above); free(f_rz); free(f_alb);
    free(refrozen_rain); free(refrozen_snow); free(snow_to_ice);
}

/* surface_physics.f90:138-153 (day, year are dead arguments there too) */
void oracle_surface_energy_and_mass_balance(oracle_state *now, const oracle_par *par,
                                            const oracle_bnd *bnd, int day, int year)
{
    int ksub;
    (void)day; (void)year;
    for (ksub = 1; ksub <= par->n_ksub; ++ksub)                                    /* :149-151 */
        oracle_energy_balance(now, par, bnd);
    oracle_mass_balance(now, par, bnd);                                            /* :152 */
}

void oracle_step(double *slab, int ld, int *mask, int nx, const oracle_par *par,
                 const oracle_bnd *bnd, unsigned char *branch)
{
    oracle_state now;
    now.nx = nx; now.ld = ld; now.slab = slab; now.mask = mask; now.branch = branch; now.rdiag = NULL;
    if (branch) memset(branch, 0, (size_t)nx);
    oracle_surface_energy_and_mass_balance(&now, par, bnd, 1, 0);
}

/* example/example.f90:88-127, optimize/run_particles.f90:124-161 */
void oracle_run(double *slab, int ld, int *mask, int nx, const oracle_par *par, const oracle_bnd *bnd,
                const double *forcing, const double *vali, int nwin, int ldf,
                int ndays, double *out, int nout, int ldo, int out_days, unsigned char *branch)
{
    oracle_run_diag(slab, ld, mask, nx, par, bnd, forcing, vali, nwin, ldf, ndays, out, nout, ldo, out_days, branch, NULL);
}

void oracle_run_diag(double *slab, int ld, int *mask, int nx, const oracle_par *par, const oracle_bnd *bnd,
                     const double *forcing, const double *vali, int nwin, int ldf,
                     int ndays, double *out, int nout, int ldo, int out_days, unsigned char *branch, double *rdiag)
{
    oracle_state now;
    int d, i;
    now.nx = nx; now.ld = ld; now.slab = slab; now.mask = mask; now.branch = NULL; now.rdiag = NULL;
    for (d = 0; d < ndays; ++d) {
        const double *fr = forcing + (size_t)(d % nwin) * 9u * (size_t)ldf;
        const double *vr = vali ? vali + (size_t)(d % nwin) * 8u * (size_t)ldf : NULL;
        /* example.f90:94-102 */
        memcpy(F(&now, OF_SF),   fr + 0 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_RF),   fr + 1 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_SWD),  fr + 2 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_LWD),  fr + 3 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_WIND), fr + 4 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_SP),   fr + 5 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_RHOA), fr + 6 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_QQ),   fr + 7 * (size_t)ldf, sizeof(double) * (size_t)nx);
        memcpy(F(&now, OF_T2M),  fr + 8 * (size_t)ldf, sizeof(double) * (size_t)nx);
        if (vr) {                                                                  /* example.f90:103-107 */
            if (bnd->tsurf) memcpy(F(&now, OF_TSURF), vr + 0 * (size_t)ldf, sizeof(double) * (size_t)nx);
            if (bnd->alb)   memcpy(F(&now, OF_ALB),   vr + 1 * (size_t)ldf, sizeof(double) * (size_t)nx);
            if (bnd->melt)  memcpy(F(&now, OF_MELT),  vr + 4 * (size_t)ldf, sizeof(double) * (size_t)nx);
            if (bnd->acc)   memcpy(F(&now, OF_ACC),   vr + 5 * (size_t)ldf, sizeof(double) * (size_t)nx);
            if (bnd->smb)   memcpy(F(&now, OF_SMB),   vr + 3 * (size_t)ldf, sizeof(double) * (size_t)nx);
        }
        if (branch) {
            now.branch = branch + (size_t)d * (size_t)nx;
            memset(now.branch, 0, (size_t)nx);
        }
        if (rdiag) now.rdiag = rdiag + (size_t)d * (size_t)nx;
        oracle_surface_energy_and_mass_balance(&now, par, bnd, d % nwin + 1, 0);   /* example.f90:112 */
        if (out && d >= ndays - out_days) {                                        /* example.f90:117-122 */
            double *o = out + (size_t)((d - (ndays - out_days)) % nout) * 8u * (size_t)ldo;
            const double *swd = fr + 2 * (size_t)ldf;
            const double *tsurf = F(&now, OF_TSURF), *alb = F(&now, OF_ALB), *smb = F(&now, OF_SMB);
            const double *melt = F(&now, OF_MELT), *acc = F(&now, OF_ACC), *shf = F(&now, OF_SHF), *lhf = F(&now, OF_LHF);
            for (i = 0; i < nx; ++i) o[0 * (size_t)ldo + i] = tsurf[i];
            for (i = 0; i < nx; ++i) o[1 * (size_t)ldo + i] = alb[i];
            for (i = 0; i < nx; ++i) o[2 * (size_t)ldo + i] = swd[i] * (1.0 - alb[i]);
            for (i = 0; i < nx; ++i) o[3 * (size_t)ldo + i] = smb[i];
            for (i = 0; i < nx; ++i) o[4 * (size_t)ldo + i] = melt[i];
            for (i = 0; i < nx; ++i) o[5 * (size_t)ldo + i] = acc[i];
            for (i = 0; i < nx; ++i) o[6 * (size_t)ldo + i] = shf[i];
            for (i = 0; i < nx; ++i) o[7 * (size_t)ldo + i] = lhf[i];
        }
    }
}

/* ---------------------------------------------------------------- averaging */

/* surface_physics.f90:601-629 */
int oracle_field_average(double *ave, const double *now, int nx, int step, double nt)
{
    int i;
    if (step == 0) { for (i = 0; i < nx; ++i) ave[i] = 0.0; }                      /* :611 */
    else if (step == 1) { for (i = 0; i < nx; ++i) ave[i] = ave[i] + now[i]; }     /* :614 */
    else if (step == 2) { for (i = 0; i < nx; ++i) ave[i] = ave[i] / nt; }         /* :621 */
    else return -1;                                                                /* :623-624 */
    return 0;
}

/* surface_physics.f90:565-597 */
int oracle_surface_physics_average(double *ave_slab, const double *now_slab, int ld, int nx, int step, double nt)
{
    static const int fields[19] = {
        OF_T2M, OF_TSURF, OF_HSNOW, OF_ALB, OF_MELT, OF_REFR, OF_SMB, OF_ACC, OF_LHF, OF_SHF, OF_LWU,   /* :574-584 */
        OF_SF, OF_RF, OF_SP, OF_LWD, OF_SWD, OF_WIND, OF_RHOA, OF_QQ                                    /* :586-593 */
    };
    int k;
    for (k = 0; k < 19; ++k) {
        int rc = oracle_field_average(ave_slab + (size_t)fields[k] * (size_t)ld,
                                      now_slab + (size_t)fields[k] * (size_t)ld, nx, step, nt);
        if (rc) return rc;
    }
    return 0;
}

/* ---------------------------------------------------------------- misfit */

/* utils.f90:23-42 */
void oracle_calculate_crmsd(const double *x1, const double *x2, int ntime, int nx, int ld, double *crmsd)
{
    int i, n;
    for (i = 0; i < nx; ++i) {
        double s, x1_avg, x2_avg, x2_std;
        s = 0.0; for (n = 0; n < ntime; ++n) s += x1[(size_t)n * (size_t)ld + i];
        x1_avg = s / ntime;                                                        /* :32 */
        s = 0.0; for (n = 0; n < ntime; ++n) s += x2[(size_t)n * (size_t)ld + i];
        x2_avg = s / ntime;                                                        /* :33 */
        s = 0.0;
        for (n = 0; n < ntime; ++n) { double d = x2[(size_t)n * (size_t)ld + i] - x2_avg; s += d * d; }
        x2_std = sqrt(s / ntime);                                                  /* :34 */
        s = 0.0;
        for (n = 0; n < ntime; ++n) {
            double d = (x1[(size_t)n * (size_t)ld + i] - x1_avg) - (x2[(size_t)n * (size_t)ld + i] - x2_avg);
            s += d * d;
        }
        crmsd[i] = sqrt(s / ntime) / x2_std;                                       /* :36 */
    }
}

/* optimize/run_particles.f90:165-175; out/vali rows are [8][ld] per day: stt alb swnet smb melt acc shf lhf */
double oracle_particle_cost_sumsq(const double *out, const double *vali, int ntime, int nx, int ld)
{
    const size_t n = (size_t)(nx > 0 ? nx : 1);
    double *c_stt = (double *)malloc(sizeof(double) * n), *c_melt = (double *)malloc(sizeof(double) * n);
    double *c_smb = (double *)malloc(sizeof(double) * n), *c_swnet = (double *)malloc(sizeof(double) * n);
    double cost = 0.0;
    int i;
    /* a day row is 8*ld doubles, so variable v of all days is a strided [ntime][8*ld] view */
    oracle_calculate_crmsd(out + 0 * (size_t)ld, vali + 0 * (size_t)ld, ntime, nx, 8 * ld, c_stt);    /* :165 */
    oracle_calculate_crmsd(out + 4 * (size_t)ld, vali + 4 * (size_t)ld, ntime, nx, 8 * ld, c_melt);   /* :166 */
    oracle_calculate_crmsd(out + 3 * (size_t)ld, vali + 3 * (size_t)ld, ntime, nx, 8 * ld, c_smb);    /* :167 */
    oracle_calculate_crmsd(out + 2 * (size_t)ld, vali + 2 * (size_t)ld, ntime, nx, 8 * ld, c_swnet);  /* :168 */
    for (i = 0; i < nx; ++i)                                                       /* :172-174 */
        cost = cost + c_stt[i] * c_stt[i] + c_melt[i] * c_melt[i] + c_smb[i] * c_smb[i] + c_swnet[i] * c_swnet[i];
    free(c_stt); free(c_melt); free(c_smb); free(c_swnet);
    return cost;
}

double oracle_particle_cost(const double *out, const double *vali, int ntime, int nx, int ld)
{
    return sqrt(oracle_particle_cost_sumsq(out, vali, ntime, nx, ld));             /* :175 */
}
