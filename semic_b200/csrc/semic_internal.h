// semic_internal.h -- types shared between the C-ABI translation unit and the two kernel
// translation units (fast / strict arithmetic).  Not part of the public boundary.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "semic_b200.h"

namespace semic {

// albedo scheme ids (string compares of surface_physics.f90:340-360 resolved once on the host)
enum { SCHEME_NONE = 0, SCHEME_SLATER = 1, SCHEME_DENBY = 2, SCHEME_ISBA = 3, SCHEME_ALEX = 4, SCHEME_UNKNOWN = 5,
       SCHEME_RUNTIME = 6 /* template value: switch on DevPar::scheme at run time */ };

// Geometry of the persistent kernel: every warp owns tiles of 32 consecutive points and a private
// shared-memory ring of kStages days fed by its own TMA bulk copies (no producer warp, no cross-warp coupling).
constexpr int kTile          = 32;                       // points per warp tile = lanes
constexpr int kStages        = 3;                        // days of forcing in flight per warp
constexpr int kLdAlign       = 1024;                     // leading dimensions: multiples of this (=> whole tiles, 16-byte rows)
constexpr int kForcingRows   = 9;
constexpr int kExtRows       = 5;                        // tsurf alb smb melt acc overrides

// Parameters in the form the kernels use.  Everything here is either a member of
// surface_param_class or a PARAMETER-ONLY expression evaluated once on the host with the
// reference's own operation order (so hoisting it changes no bits).
struct DevPar {
    double ceff, albi, albl, alb_smax, alb_smin, hcrit, rcrit, amp, csh, clh, tmin, tmax,
           tstic, tsticsub, tau_a, tau_f, w_crit, mcrit, afac, tmid;
    // exact hoists (same operations on the same operands as the reference performs per point)
    double csh_cap;          // csh*cap                      (:387 prefix)
    double rhow_clm;         // rhow*clm                     (:266,:286)
    double one_m_frz;        // 1.0 - f_rz, f_rz = rcrit     (:284,:298)
    double hcrit_eps;        // hcrit + epsil                (:338)
    double slater_f;         // 1.0/(t0 - tmin)              (:518)
    double smax_m_smin;      // alb_smax - alb_smin          (:487,:529,:541,:361)
    double isba_dry;         // tau_a*tstic/tau, tau = tstic (:484)
    double isba_wet;         // dexp(-tau_f*tstic/tau)       (:486)  host libm, like the reference
    double isba_wcr;         // w_crn/rhow                   (:487)
    // tolerance-funded reciprocals (fast arithmetic only; <= 1 ulp each)
    double sub_over_ceff;    // tsticsub/ceff
    double ceff_over_sub;    // ceff/tsticsub
    double tstic_over_ceff;  // tstic/ceff
    double ceff_over_tstic;  // ceff/tstic
    double inv_tstic;        // 1/tstic
    double inv_rhow_clm;     // 1/(rhow*clm)
    double inv_hcrit_eps;    // 1/(hcrit+epsil)
    double inv_amp;          // 1/amp
    double inv_amp2;         // 1/(amp*amp)
    double inv_mcrit;        // 1/mcrit
    double isba_new;         // tstic/(w_crn/rhow)*(alb_smax-alb_smin)
    double frz_rhow_clm;     // (1-f_rz)*rhow*clm
    double rcrit_zero;       // rcrit * (+0): the snow-refreezing term of a day without melt (keeps the sign of rcrit)
    double melt_per_kelvin;  // (ceff/tstic) * 1/(rhow*clm): melt or refreezing rate per kelvin of the diurnal integral (:254-266, :257-286)
    double inv_tsticsub;     // 1/tsticsub, 1/ceff: correctly rounded reciprocals for the exact quotients of :202, :249
    double inv_ceff;
    int    n_ksub;
    int    scheme;
};

// Series (forcing, validation, daily output) are stored TILE-INTERLEAVED in HBM:
//   element (day d, var v, point i) at  base[((d*ntl + i/32)*nvar + v)*32 + i%32],  ntl = ld/32,
// so the nvar rows a warp needs for one day are ONE contiguous block (9*256 B of forcing = one TMA bulk copy,
// 8*256 B of output = eight stores off one base pointer).  The state slab stays SoA (its pinned host mirror
// is what Fortran/Python array components alias).
struct RunArgs {
    // state slab (device): field f of point i at slab[f*ld + i]
    double       *slab;
    const int    *mask;
    int           nx;
    int           ld;
    int           sld;          // leading dimension of the tiled series (forcing, vali, out); = ld except for point slices
    // forcing: tiled series (f_tiled=1, day d -> window row (f_day0+d) % f_nwin), or the slab's own SoA forcing
    // fields (f_tiled=0: variable v of point i at forcing[f_var_off[v] + i]; the drop-in single-day call)
    const double *forcing;
    int           f_tiled;
    long long     f_var_off[kForcingRows];
    int           f_nwin;
    int           f_day0;
    // external override fields (tiled validation series, 8 rows per day) or NULL
    const double *vali;
    // daily output ring (tiled, o_nvar rows per day) or NULL: day d of the run goes to row (d - out_start) % o_nring
    double       *out;
    int           o_nring;
    int           o_nvar;       // 8, or 9 with the branch bitmap
    int           out_start;    // first day (0-based within this run) that is written
    int           avg_days;     // > 1: rows of `out` are means over periods of avg_days days (k_run_lean only)
    int           ndays;
    int           eb_steps;     // sub-steps of energy_balance per day (n_ksub; 1 for energy_balance alone; 0 = skip)
    int           do_mb;        // run mass_balance
    DevPar        P;
    semic_bnd     B;
};

// ensemble (config 5): every particle starts from the same initial slab
struct EnsArgs {
    const double *slab0;        // initial state [31][ld]
    const int    *mask;
    int           nx, ld;
    const double *forcing;      // tiled [ntime][ld/32][9][32]
    const double *vali;         // tiled [ntime][ld/32][8][32]
    const double *inv_var;      // [4][ld]  1/var of vali stt, melt, smb, swnet (zero variance -> inf, like utils.f90:36)
    const DevPar *pars;         // [npar]
    int           npar;
    int           ntime, nloop;
    double       *partial;      // [npar][ntiles] per-CTA partial sums (deterministic two-stage reduce)
    int           ntiles;       // ceil(nx/32)
    int           uni_scheme;   // the scheme id all particles share, or -1
    int           uni_ksub;     // the n_ksub all particles share, or 0
};

// Peer-memory exchange of the ensemble's per-particle sums (NVLink / NVSwitch, one process per GPU).  Every rank owns a buffer
// [2 flags][2 slots][kP2PMaxPar doubles] that all ranks have mapped (CUDA IPC); slot = epoch & 1.
constexpr int kP2PMaxPar = 65536;
struct P2PArgs {
    double *const *data;                    // device array [nranks]: base of rank r's two slots
    unsigned long long *const *flags;       // device array [nranks]: rank r's two epoch flags
    unsigned *done;                         // this rank's CTA counter (zero between launches)
    int *err;                               // set to 1 if a peer did not arrive in time
    int nranks, rank;
    unsigned long long epoch;               // > 0, the same on every rank
    long long timeout_cycles;
};

// one-time per-device set-up of the fast arithmetic (the 2^(j/2048) table); cheap no-op afterwards.  The launchers call it
// themselves; timed entry points call it before their first event so that it never lands inside a timed region
cudaError_t prepare_fast();
// implemented once per arithmetic mode (semic_kernels_fast.cu / semic_kernels_strict.cu)
cudaError_t launch_run_fast(const RunArgs &a, bool generic, cudaStream_t st, int *launches);
cudaError_t launch_run_strict(const RunArgs &a, bool generic, cudaStream_t st, int *launches);
// inv_var and partial are scratch buffers owned by the caller; sumsq_dev receives npar doubles
cudaError_t launch_vali_invvar(const EnsArgs &a, double *inv_var, cudaStream_t st);
// p2p != NULL: the second reduction stage also sums over the ranks through peer memory (k_ens_reduce_p2p) and sumsq_dev
// receives the GLOBAL sums; mid (may be NULL) is recorded between the trajectory kernel and the reduction
cudaError_t launch_ensemble_fast(const EnsArgs &a, double *sumsq_dev, cudaStream_t st, const P2PArgs *p2p = nullptr,
                                 cudaEvent_t mid = nullptr);
cudaError_t launch_elementals_fast(int which, const double *const *in, const int *mask, const double *sc,
                                   double *const *out, int n, cudaStream_t st);

}  // namespace semic
