/*
 * semic_b200.h -- C-ABI of the B200-native SEMIC time step (libsemic_b200.so).
 *
 * Drop-in boundary for ONE path of mkrapp/semic: the public procedures of
 * `module surface_physics` (surface_physics.f90:121-130) and the ensemble misfit of
 * optimize/run_particles.f90.  Plain pointers and sizes only; every entry returns an int
 * status (0 = ok, <0 = error, text via semic_last_error()).  The Fortran shim
 * (fortran/surface_physics_b200.f90, ISO_C_BINDING) and the Python module
 * (semic_b200/surface_physics.py, ctypes) bind exactly these symbols.
 *
 * There is no CPU fallback: anything that computes needs a CUDA device and fails with
 * SEMIC_ERR_CUDA otherwise.  Host-only entries (par_load, boundary_define, print_*) work anywhere.
 */
#ifndef SEMIC_B200_H
#define SEMIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEMIC_STRLEN   256   /* character(len=256), surface_physics.f90:29-31 */
#define SEMIC_NBOUNDARY 30   /* boundary(30), surface_physics.f90:30 */

enum {
    SEMIC_OK = 0,
    SEMIC_ERR_ARG = -1,      /* bad argument                                   */
    SEMIC_ERR_CUDA = -2,     /* CUDA runtime error (incl. "no device")          */
    SEMIC_ERR_IO = -3,       /* namelist file missing / unreadable              */
    SEMIC_ERR_PARSE = -4,    /* namelist syntax error, group not found          */
    SEMIC_ERR_STEP = -5,     /* surface_physics_average: step not recognized / nt missing (:616-624) */
    SEMIC_ERR_NCCL = -6
};

/* field ids = declaration order of surface_state_class, surface_physics.f90:57-89 */
enum {
    SEMIC_F_T2M = 0, SEMIC_F_TSURF, SEMIC_F_HSNOW, SEMIC_F_HICE, SEMIC_F_ALB, SEMIC_F_ALB_SNOW,
    SEMIC_F_MELT, SEMIC_F_MELTED_SNOW, SEMIC_F_MELTED_ICE, SEMIC_F_REFR, SEMIC_F_SMB, SEMIC_F_ACC,
    SEMIC_F_LHF, SEMIC_F_SHF, SEMIC_F_LWU, SEMIC_F_SUBL, SEMIC_F_EVAP, SEMIC_F_SMB_SNOW,
    SEMIC_F_SMB_ICE, SEMIC_F_RUNOFF, SEMIC_F_QMR, SEMIC_F_QMR_RES, SEMIC_F_AMP,
    SEMIC_F_SF, SEMIC_F_RF, SEMIC_F_SP, SEMIC_F_LWD, SEMIC_F_SWD, SEMIC_F_WIND, SEMIC_F_RHOA, SEMIC_F_QQ,
    SEMIC_NFIELDS,             /* 31 double fields */
    SEMIC_F_MASK = SEMIC_NFIELDS   /* the integer mask, bit 31 of a field set */
};
#define SEMIC_FIELDS_ALL 0xFFFFFFFFull

/* forcing series variable order = column order of the forcing files, utils.f90:63-67 */
enum { SEMIC_V_SF = 0, SEMIC_V_RF, SEMIC_V_SWD, SEMIC_V_LWD, SEMIC_V_WIND, SEMIC_V_SP, SEMIC_V_RHOA,
       SEMIC_V_QQ, SEMIC_V_T2M, SEMIC_NFORCING };
/* validation / daily-output variable order, utils.f90:89-93 and example/example.f90:117-122 */
enum { SEMIC_O_STT = 0, SEMIC_O_ALB, SEMIC_O_SWNET, SEMIC_O_SMB, SEMIC_O_MELT, SEMIC_O_ACC, SEMIC_O_SHF,
       SEMIC_O_LHF, SEMIC_NOUT, SEMIC_O_BRANCH = SEMIC_NOUT /* optional 9th row: branch bitmap as a double */ };

/* branch bitmap bits (threshold decisions of a point-day; the flip report compares them with the oracle's) */
enum { SEMIC_B_TS_LT_T0 = 1, SEMIC_B_CLAMP = 2, SEMIC_B_GATE = 4, SEMIC_B_DIURNAL_LO = 8, SEMIC_B_DIURNAL_HI = 16,
       SEMIC_B_MELT_POS = 32, SEMIC_B_HSNOW_ZERO = 64, SEMIC_B_SNOW2ICE = 128 };

/* surface_param_class, surface_physics.f90:28-54.  Strings are blank-padded or NUL-terminated. */
typedef struct semic_par {
    char   name[SEMIC_STRLEN];
    char   boundary[SEMIC_NBOUNDARY][SEMIC_STRLEN];
    char   alb_scheme[SEMIC_STRLEN];
    int    nx;
    int    n_ksub;
    double ceff, albi, albl, alb_smax, alb_smin, hcrit, rcrit, amp, csh, clh, tmin, tmax,
           tstic, tsticsub, tau_a, tau_f, w_crit, mcrit, afac, tmid;
} semic_par;

/* boundary_opt_class, surface_physics.f90:92-105 (logical -> int 0/1) */
typedef struct semic_bnd {
    int t2m, tsurf, hsnow, alb, melt, refr, smb, acc, lhf, shf, subl, amp;
} semic_bnd;

typedef struct semic_state  semic_state;    /* surface_state_class: device SoA + pinned host mirror */
typedef struct semic_series semic_series;   /* device-resident time series, logically [day][var][point]; stored
                                               tile-interleaved in HBM ([day][point/32][var][32]), see DESIGN.md */
typedef struct semic_comm   semic_comm;     /* NCCL communicator wrapper (ensemble all-reduce)        */

/* ---- module procedures (same order as surface_physics.f90:121-130) ------------------------- */

/* surface_alloc(now,npts), surface_physics.f90:633-677.  Unlike the reference, contents are zero-filled. */
int semic_surface_alloc(semic_state **now, int npts);
/* surface_dealloc(now), surface_physics.f90:681-725 */
int semic_surface_dealloc(semic_state *now);
/* surface_physics_par_load(par,filename), surface_physics.f90:729-814: reads group &surface_physics;
 * keys not present keep their incoming value; tsticsub = tstic/dble(n_ksub) (:810). */
int semic_surface_physics_par_load(semic_par *par, const char *filename);
/* surface_boundary_define(bnd,boundary), surface_physics.f90:819-878.  `boundary` is n strings of `len`
 * bytes each (blank padded or NUL terminated); unknown names are ignored (:871-872). */
int semic_surface_boundary_define(semic_bnd *bnd, const char *boundary, int n, int len);
/* surface_energy_and_mass_balance(now,par,bnd,day,year), surface_physics.f90:138-153.
 * Drop-in semantics: uploads the host mirror (fields in the state's upload set), runs one fused
 * kernel for n_ksub x energy_balance + mass_balance, downloads the download set, synchronises. */
int semic_surface_energy_and_mass_balance(semic_state *now, const semic_par *par, const semic_bnd *bnd,
                                          int day, int year);
/* energy_balance / mass_balance alone, surface_physics.f90:158-214 and :232-374 (public in the module) */
int semic_energy_balance(semic_state *now, const semic_par *par, const semic_bnd *bnd, int day, int year);
int semic_mass_balance(semic_state *now, const semic_par *par, const semic_bnd *bnd, int day, int year);
/* surface_physics_average(ave,now,step,nt), surface_physics.f90:565-597; step = "init" | "step" | "end";
 * nt < 0 means "not present".  Device-side on the two states' slabs (no host traffic). */
int semic_surface_physics_average(semic_state *ave, const semic_state *now, const char *step, double nt);
/* print_param / print_boundary_opt, surface_physics.f90:882-932 (stdout) */
int semic_print_param(const semic_par *par);
int semic_print_boundary_opt(const semic_bnd *bnd);
/* public elementals, surface_physics.f90:378-438, evaluated on the device over n values */
int semic_sensible_heat_flux(const double *ts, const double *ta, const double *wind, const double *rhoa,
                             double csh, double cap, double *shf, int n);
int semic_latent_heat_flux(const double *ts, const double *wind, const double *shum, const double *sp,
                           const double *rhoatm, const int *mask, double clh, double eps, double cls, double clv,
                           double *lhf, double *subl, double *evap, int n);
int semic_longwave_upward(const double *ts, double sigma, double *lwout, int n);
/* One value of the same, as a function RESULT (NaN when the call fails): what a Fortran ELEMENTAL specific can forward to through
 * a PURE bind(C) interface with VALUE dummies -- no result reaches Fortran through a pointer, so an optimising compiler cannot
 * assume the call left it untouched.  which: 0 = lhf, 1 = subl, 2 = evap (surface_physics.f90:393-430). */
double semic_sensible_heat_flux_1(double ts, double ta, double wind, double rhoa, double csh, double cap);
double semic_latent_heat_flux_1(double ts, double wind, double shum, double sp, double rhoatm, int mask,
                                double clh, double eps, double cls, double clv, int which);
double semic_longwave_upward_1(double ts, double sigma);
/* public constants sigm, cap, eps, cls, clv (surface_physics.f90:129) by name; NaN if unknown */
double semic_constant(const char *name);

/* ---- state access ---------------------------------------------------------------------------- */
int     semic_state_npts(const semic_state *now);
int     semic_state_ld(const semic_state *now);                 /* leading dimension of the slabs */
double *semic_state_field(semic_state *now, int field);          /* pinned host mirror, npts doubles */
int    *semic_state_mask(semic_state *now);                      /* pinned host mirror, npts ints    */
double *semic_state_device_slab(semic_state *now);               /* device [31][ld] */
int     semic_state_upload(semic_state *now, uint64_t fields);   /* host mirror -> device, bit f = field f, bit 31 = mask */
int     semic_state_download(semic_state *now, uint64_t fields); /* device -> host mirror */
/* which fields the drop-in step moves each call (default: everything both ways) */
int     semic_state_set_sync(semic_state *now, uint64_t upload_fields, uint64_t download_fields);

/* ---- device-resident series -------------------------------------------------------------------- */
int     semic_series_alloc(semic_series **s, int npts, int nvar, int ndays);
int     semic_series_free(semic_series *s);
int     semic_series_npts(const semic_series *s);
int     semic_series_ld(const semic_series *s);
int     semic_series_nvar(const semic_series *s);
int     semic_series_ndays(const semic_series *s);
double *semic_series_device_ptr(semic_series *s);
/* host layout [ndays][nvar][host_ld]; copies points [0,npts) of days [day0, day0+ndays) */
int     semic_series_upload(semic_series *s, int day0, int ndays, const double *host, int host_ld);
int     semic_series_download(const semic_series *s, int day0, int ndays, double *host, int host_ld,
                              int point0, int npoints);
/* synthetic forcing of SURVEY.md 8(d) generated on the device (counter-based splitmix64): fills all
 * ndays x 9 rows for global points [point0, point0+npts); row d holds calendar day dayofs + d*daystride
 * (daystride > 1 samples the seasonal cycle into a short periodic window). */
int     semic_series_synth_forcing(semic_series *s, uint64_t seed, int64_t point0, int dayofs, int daystride);
/* synthetic mask: fractions of ice(2)/land(1), rest ocean(0); writes the host mirror AND the device */
int     semic_state_synth_mask(semic_state *now, uint64_t seed, int64_t point0, double frac_ice, double frac_land);
/* fill one field with a constant on host mirror and device (initial conditions) */
int     semic_state_fill(semic_state *now, int field, double value);

/* ---- multi-day run (persistent kernel) ------------------------------------------------------------ */
typedef struct semic_run_opts {
    int strict;        /* 0 = fast math (default), 1 = reference operation order, IEEE div, no FMA contraction */
    int out_days;      /* write daily output for the LAST out_days days of the run (0 = none)       */
    int reserved0;     /* unused, keep 0 (keeps the layout of round 1)                                */
    int avg_days;      /* > 1: the rows of `out` are the MEANS of the 8 daily fields over consecutive periods
                          of avg_days days of the last out_days days (last period may be shorter) -- the protocol of
                          surface_physics_average (surface_physics.f90:565-629: init / step / end) kept in registers.
                          Plain runs only (no boundary overrides, 8-row output, fast arithmetic). 0 or 1 = daily rows */
    int reserved[4];
} semic_run_opts;

/* Runs ndays days: day d uses forcing row (day0 + d) % ndays(forcing).  `vali` (8 vars, same window)
 * supplies the external fields for bnd%tsurf/alb/melt/acc/smb as example/example.f90:103-107 does; NULL =
 * flagged fields keep their state value.  `out` (8 or 9 vars) is a ring: day d of the last out_days days
 * goes to row (d - (ndays-out_days)) % ndays(out).  State is read from and written back to the DEVICE
 * slab of `now` (all 31 fields + final forcing, as after the reference's last call).
 * kernel_ms: CUDA-event time of the kernel on the state's stream (may be NULL). */
int semic_run(semic_state *now, const semic_par *par, const semic_bnd *bnd, const semic_series *forcing,
              const semic_series *vali, int day0, int ndays, semic_series *out, const semic_run_opts *opts,
              float *kernel_ms);

/* End-to-end variant with HOST buffers: forcing [nwin][9][ldh] (and vali [nwin][8][ldh] or NULL) are streamed
 * host->device in chunks of chunk_days, overlapped with the kernel; the daily output of the last out_days days
 * is streamed back to h_out [out_days][8][ldh].  Host buffers should be pinned (semic_host_alloc).
 * opts->avg_days > 1 (plain fast-mode runs): h_out holds ceil(out_days/avg_days) rows, the means over consecutive periods
 * (surface_physics_average protocol); needs chunk_days % avg_days == 0 and (ndays - out_days) % chunk_days == 0. */
int semic_run_host(semic_state *now, const semic_par *par, const semic_bnd *bnd, const double *h_forcing,
                   const double *h_vali, int nwin, int ldh, int ndays, double *h_out, int out_days,
                   int chunk_days, const semic_run_opts *opts, float *total_ms);
void *semic_host_alloc(size_t bytes);     /* pinned host memory */
int   semic_host_free(void *p);

/* ---- ensemble misfit (optimize/run_particles.f90:104-178, utils.f90:23-42) ------------------------ */
/* For each of npar parameter sets: start from the state in `init` (not modified), run nloop x ntime days on
 * forcing/vali (ntime-day windows), and accumulate over the local points
 *   sumsq[p] = sum_n crmsd_stt(n)^2 + crmsd_melt(n)^2 + crmsd_smb(n)^2 + crmsd_swnet(n)^2   (:172-174)
 * from the last loop's daily fields.  sumsq is a HOST array of npar doubles (cost = sqrt after the all-reduce).
 * If comm != NULL the per-particle sums are all-reduced (ncclSum) over the communicator first. */
int semic_ensemble_cost(const semic_state *init, const semic_par *pars, int npar, const semic_bnd *bnd,
                        const semic_series *forcing, const semic_series *vali, int ntime, int nloop,
                        semic_comm *comm, double *sumsq, float *kernel_ms);

/* ---- NCCL plumbing -------------------------------------------------------------------------------- */
int semic_comm_unique_id(char id[128]);                                   /* ncclGetUniqueId */
int semic_comm_init(semic_comm **c, const char id[128], int nranks, int rank);   /* ncclCommInitRank on the current device */
int semic_comm_allreduce_sum(semic_comm *c, double *dev_buf, int count);  /* in place, on the comm's stream, synchronises */
/* 1 if the communicator exchanges through NVLink peer memory (CUDA IPC mappings of every rank's buffer; the all-reduce of
 * semic_ensemble_cost is then part of its reduction kernel), 0 if through ncclAllReduce; `why` receives the reason for 0.
 * Env SEMIC_COMM_P2P=0 at semic_comm_init keeps NCCL (all ranks alike). */
int semic_comm_p2p(const semic_comm *c, char *why, int why_len);
int semic_comm_last_allreduce_ms(const semic_comm *c, float *ms);         /* device time (CUDA events on the comm's stream) of the last all-reduce */
int semic_comm_free(semic_comm *c);

/* ---- misc -------------------------------------------------------------------------------------------- */
int         semic_set_device(int device);
int         semic_device_count(void);
int         semic_device_synchronize(void);
const char *semic_last_error(void);
const char *semic_version(void);
/* number of time-step kernel launches this process has made (bench: gpu_launches) */
int         semic_kernel_launches(void);
/* measured-peak helpers for the roofline (bench only): DFMA and copy micro-kernels, return per-launch ms */
int semic_probe_dfma(int iters, float *ms, double *flops);
int semic_probe_copy(size_t bytes, float *ms);
/* writes a buffer larger than L2 (flushes the 126 MB L2 between timed iterations) */
int semic_flush_l2(void);

#ifdef __cplusplus
}
#endif
#endif /* SEMIC_B200_H */
