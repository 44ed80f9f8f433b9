/*
 * semic_oracle.h -- CPU restatement of SEMIC's daily time step (TEST INFRASTRUCTURE ONLY).
 *
 * This is the checker for the CUDA path in semic_b200/csrc.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (libsemic_b200.so) never links, loads or calls anything in this directory.
 *
 * PARITY PINNED (round 2): the reference (mkrapp/semic) is Fortran and the image has no Fortran
 * compiler, so oracle/f90_to_c.py -- a physics-free source-to-source translator -- turns the
 * reference's own surface_physics.f90, utils.f90, example.f90 and run_particles.f90 into C at build
 * time (`make -C oracle ref` -> oracle/_ref/, git-ignored).  tests/test_ref_pin.py demands bit-for-bit
 * equality between this restatement and that build over 6 schemes x 10 boundary-flag sets x the three
 * shipped data sets, the corner cases, the elementals, CRMSD, averaging and both translated programs;
 * tests/golden/ref_*.npy are outputs of the translated reference programs on the shipped data
 * (tests/golden/make_ref_golden.py) and are reproduced bit for bit (tests/test_oracle.py).  What stays
 * unpinned: a gfortran build's libm calls and instruction selection, which the translation shares with
 * this file by construction (same gcc, same glibc).
 *
 * Every loop below is one Fortran whole-array statement (the reference is written in array
 * syntax; gfortran emits one sweep over nx per statement and heap-allocates the automatic
 * arrays on every call), so this file doubles as the CPU baseline with the reference's own
 * memory behaviour.  Build: gcc -O3 -fno-fast-math -ffp-contract=off (x86-64 SSE2 doubles,
 * no FMA, no reassociation: the semantics of the reference's gfortran -O2/-O3 build).
 */
#ifndef SEMIC_ORACLE_H
#define SEMIC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* field order of surface_state_class, surface_physics.f90:56-90 */
enum {
    OF_T2M = 0, OF_TSURF, OF_HSNOW, OF_HICE, OF_ALB, OF_ALB_SNOW, OF_MELT, OF_MELTED_SNOW,
    OF_MELTED_ICE, OF_REFR, OF_SMB, OF_ACC, OF_LHF, OF_SHF, OF_LWU, OF_SUBL, OF_EVAP,
    OF_SMB_SNOW, OF_SMB_ICE, OF_RUNOFF, OF_QMR, OF_QMR_RES, OF_AMP,
    OF_SF, OF_RF, OF_SP, OF_LWD, OF_SWD, OF_WIND, OF_RHOA, OF_QQ,
    OF_NFIELDS
};

/* surface_param_class, surface_physics.f90:28-54 (numeric part + scheme name) */
typedef struct oracle_par {
    char   alb_scheme[256];
    int    nx;
    int    n_ksub;
    double ceff, albi, albl, alb_smax, alb_smin, hcrit, rcrit, amp, csh, clh, tmin, tmax,
           tstic, tsticsub, tau_a, tau_f, w_crit, mcrit, afac, tmid;
} oracle_par;

/* boundary_opt_class, surface_physics.f90:92-105 */
typedef struct oracle_bnd {
    int t2m, tsurf, hsnow, alb, melt, refr, smb, acc, lhf, shf, subl, amp;
} oracle_bnd;

/* state: slab[f*ld + i] is field f (enum above) of point i; mask[i] in {0,1,2} */
typedef struct oracle_state {
    int     nx;
    int     ld;
    double *slab;
    int    *mask;
    unsigned char *branch;   /* optional [nx] per-day branch bitmap, NULL = off */
    double *rdiag;           /* optional [nx] per-day tmean/amp of the diurnal cycle (conditioning report), NULL = off */
} oracle_state;

/* branch bitmap bits (threshold decisions of one day, used for the flip report) */
enum {
    OB_TS_LT_T0   = 1,   /* ts < t0 in latent_heat_flux, last sub-step (:415)          */
    OB_CLAMP      = 2,   /* tsurf > t0 clamp fired, last sub-step (:201)                */
    OB_GATE       = 4,   /* mask==2 .or. hsnow>0 (:201,:209)                            */
    OB_DIURNAL_LO = 8,   /* diurnal branch id bit0: 0 = tmean+amp<0, 1 = interior, 2 = above only */
    OB_DIURNAL_HI = 16,
    OB_MELT_POS   = 32,  /* melt > 0 after :264-280                                     */
    OB_HSNOW_ZERO = 64,  /* hsnow == 0 after :315-324                                   */
    OB_SNOW2ICE   = 128  /* snow_to_ice > 0 (:323)                                      */
};

/* elementals (surface_physics.f90:378-560) */
double oracle_sensible_heat_flux(double ts, double ta, double wind, double rhoa, double csh, double cap);
void   oracle_latent_heat_flux(double ts, double wind, double shum, double sp, double rhoatm, int mask,
                               double clh, double eps, double cls, double clv,
                               double *lhf, double *subl, double *evap);
double oracle_longwave_upward(double ts, double sigma);
void   oracle_diurnal_cycle(double amp, double tmean, double *above, double *below);
void   oracle_albedo_isba(double *alb, double sf, double melt, double tstic, double tau, double tau_a,
                          double tau_f, double w_crn, double mcrit, double alb_smin, double alb_smax);
double oracle_albedo_slater(double tsurf, double tmin, double tmax, double alb_smax, double alb_smin);
double oracle_albedo_denby(double melt, double alb_smax, double alb_smin, double mcrit);
double oracle_ew_sat(double t);
double oracle_ei_sat(double t);

/* time step (surface_physics.f90:138-374) */
void oracle_energy_balance(oracle_state *now, const oracle_par *par, const oracle_bnd *bnd);
void oracle_mass_balance(oracle_state *now, const oracle_par *par, const oracle_bnd *bnd);
void oracle_surface_energy_and_mass_balance(oracle_state *now, const oracle_par *par,
                                            const oracle_bnd *bnd, int day, int year);

/* flat entry for ctypes: one day on a slab */
void oracle_step(double *slab, int ld, int *mask, int nx, const oracle_par *par,
                 const oracle_bnd *bnd, unsigned char *branch);

/*
 * Day loop of example/example.f90:88-127 and optimize/run_particles.f90:124-161.
 *   forcing [nwin][9][ldf]  columns sf rf swd lwd wind sp rhoa qq t2m (utils.f90:63-67), day d uses row d % nwin
 *   vali    [nwin][8][ldf]  columns stt alb swnet smb melt acc shf lhf (utils.f90:89-93) or NULL;
 *           overrides tsurf/alb/melt/acc/smb when the bnd flag is set (example.f90:103-107)
 *   out     [nout][8][ldo]  tsurf alb swnet smb melt acc shf lhf (example.f90:117-122); day d of the
 *           LAST out_days days goes to row (d - (ndays-out_days)) % nout; NULL = none
 *   branch  [ndays][nx] or NULL
 */
void oracle_run(double *slab, int ld, int *mask, int nx, const oracle_par *par, const oracle_bnd *bnd,
                const double *forcing, const double *vali, int nwin, int ldf,
                int ndays, double *out, int nout, int ldo, int out_days, unsigned char *branch);

/* the same, plus rdiag [ndays][nx] (or NULL): r = tmean/amp of every point-day, for the conditioning report */
void oracle_run_diag(double *slab, int ld, int *mask, int nx, const oracle_par *par, const oracle_bnd *bnd,
                     const double *forcing, const double *vali, int nwin, int ldf,
                     int ndays, double *out, int nout, int ldo, int out_days, unsigned char *branch, double *rdiag);

/* surface_physics.f90:601-629; step: 0 "init", 1 "step", 2 "end"; returns -1 for anything else */
int oracle_field_average(double *ave, const double *now, int nx, int step, double nt);
/* surface_physics.f90:565-597: the 19 averaged fields of two slabs */
int oracle_surface_physics_average(double *ave_slab, const double *now_slab, int ld, int nx, int step, double nt);

/* utils.f90:23-42; x1,x2 are [ntime][ld] (point-contiguous), crmsd[nx] */
void oracle_calculate_crmsd(const double *x1, const double *x2, int ntime, int nx, int ld, double *crmsd);

/* optimize/run_particles.f90:165-175 on model out [ntime][8][ld] and vali [ntime][8][ld] */
double oracle_particle_cost(const double *out, const double *vali, int ntime, int nx, int ld);
/* same, but returns the sum of squares before the sqrt (per-shard partial of the all-reduce) */
double oracle_particle_cost_sumsq(const double *out, const double *vali, int ntime, int nx, int ld);

#ifdef __cplusplus
}
#endif
#endif
