// semic_capi.cu -- the C-ABI of include/semic_b200.h: handles, device SoA mirrors, the drop-in
// daily step, the multi-day run, the host-streamed run, series helpers, synthetic forcing,
// averaging, NCCL plumbing.  The physics kernels live in semic_kernels.cuh (two arithmetic modes).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <mutex>
#include <vector>

#include "semic_b200.h"
#include "semic_internal.h"

namespace semic {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static int cuda_fail(cudaError_t e, const char *what)
{
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return SEMIC_ERR_CUDA;
}
#define CU(call)                                                    \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);       \
    } while (0)

int namelist_load(semic_par *par, const char *filename);

// ------------------------------------------------------------------------------------------------
// strings (blank padded Fortran style or NUL terminated)
// ------------------------------------------------------------------------------------------------
static std::string ftrim(const char *s, int len)
{
    int n = 0;
    while (n < len && s[n] != '\0') ++n;
    while (n > 0 && s[n - 1] == ' ') --n;
    return std::string(s, (size_t)n);
}

static int scheme_id(const semic_par *par)
{
    const std::string s = ftrim(par->alb_scheme, SEMIC_STRLEN);    // trim(par%alb_scheme), :340-360
    if (s == "slater") return SCHEME_SLATER;
    if (s == "denby") return SCHEME_DENBY;
    if (s == "isba") return SCHEME_ISBA;
    if (s == "none") return SCHEME_NONE;
    if (s == "alex") return SCHEME_ALEX;
    return SCHEME_UNKNOWN;                                         // Q7: alb_snow stays unchanged
}

static DevPar make_devpar(const semic_par *par)
{
    const double t0 = 273.15, cap = 1000.0, rhow = 1000.0, clm = 3.30e5, epsil = 2.220446049250313e-16;
    DevPar P;
    std::memset(&P, 0, sizeof P);
    P.ceff = par->ceff; P.albi = par->albi; P.albl = par->albl; P.alb_smax = par->alb_smax; P.alb_smin = par->alb_smin;
    P.hcrit = par->hcrit; P.rcrit = par->rcrit; P.amp = par->amp; P.csh = par->csh; P.clh = par->clh;
    P.tmin = par->tmin; P.tmax = par->tmax; P.tstic = par->tstic; P.tsticsub = par->tsticsub;
    P.tau_a = par->tau_a; P.tau_f = par->tau_f; P.w_crit = par->w_crit; P.mcrit = par->mcrit;
    P.afac = par->afac; P.tmid = par->tmid;
    // parameter-only expressions in the reference's operation order (volatile blocks contraction)
    volatile double v;
    v = par->csh * cap;                         P.csh_cap = v;
    v = rhow * clm;                             P.rhow_clm = v;
    v = 1.0 - par->rcrit;                       P.one_m_frz = v;
    v = par->hcrit + epsil;                     P.hcrit_eps = v;
    v = t0 - par->tmin; v = 1.0 / v;            P.slater_f = v;
    v = par->alb_smax - par->alb_smin;          P.smax_m_smin = v;
    v = par->tau_a * par->tstic; v = v / par->tstic;            P.isba_dry = v;
    v = -par->tau_f * par->tstic; v = v / par->tstic; v = std::exp(v);   P.isba_wet = v;
    v = par->w_crit / rhow;                     P.isba_wcr = v;
    P.sub_over_ceff = par->tsticsub / par->ceff;
    P.ceff_over_sub = par->ceff / par->tsticsub;
    P.tstic_over_ceff = par->tstic / par->ceff;
    P.ceff_over_tstic = par->ceff / par->tstic;
    P.inv_tstic = 1.0 / par->tstic;
    P.inv_rhow_clm = 1.0 / P.rhow_clm;
    P.inv_hcrit_eps = 1.0 / P.hcrit_eps;
    P.inv_amp = 1.0 / par->amp;
    P.inv_amp2 = 1.0 / (par->amp * par->amp);
    P.inv_mcrit = 1.0 / par->mcrit;
    P.isba_new = par->tstic / P.isba_wcr * P.smax_m_smin;
    P.frz_rhow_clm = P.one_m_frz * rhow * clm;
    P.melt_per_kelvin = P.ceff_over_tstic * P.inv_rhow_clm;
    { volatile double z = 0.0; P.rcrit_zero = par->rcrit * z; }
    P.inv_tsticsub = 1.0 / par->tsticsub;
    P.inv_ceff = 1.0 / par->ceff;
    P.n_ksub = par->n_ksub;
    P.scheme = scheme_id(par);
    return P;
}

static bool any_consulted_flag(const semic_bnd *b)
{
    // flags the time step reads; t2m, hsnow, subl are never consulted
    return b->shf || b->lhf || b->tsurf || b->amp || b->melt || b->refr || b->acc || b->smb || b->alb;
}

}  // namespace semic

using namespace semic;

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct semic_state {
    int npts = 0, ld = 0, device = 0;
    double *d_slab = nullptr, *h_slab = nullptr;
    int *d_mask = nullptr, *h_mask = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t up_fields = SEMIC_FIELDS_ALL, down_fields = 0x7FFFFFull;   // download: the 23 non-forcing fields
    // created on first use by the sliced per-day call (large grids): copy-in / copy-out streams and per-slice events
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_done;
};

struct semic_series {
    int npts = 0, ld = 0, nvar = 0, ndays = 0, device = 0;
    double *d = nullptr;
};

struct semic_comm {
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;      // around the last all-reduce, on `stream`
    float last_ms = 0.f;
    int nranks = 1, rank = 0;
    // peer-memory exchange (NVLink): every rank's buffer [2 flags | 2 x kP2PMaxPar doubles] mapped on every rank through CUDA IPC
    bool p2p_ready = false;
    char p2p_why[160] = "";                      // why the peer-memory path is off (then the all-reduce is NCCL's)
    void *p2p_own = nullptr;                     // this rank's buffer
    std::vector<void *> p2p_peer;                // mapped peers (own entry = p2p_own)
    double **d_data = nullptr;                   // device arrays of the per-rank pointers
    unsigned long long **d_flags = nullptr;
    unsigned *d_done = nullptr;
    int *d_err = nullptr;
    unsigned long long epoch = 0;
    int device = 0;
};

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

namespace semic {
DevPar make_devpar_public(const semic_par *par) { return make_devpar(par); }
bool any_consulted_flag_public(const semic_bnd *b) { return any_consulted_flag(b); }
int state_info(const semic_state *s, int *npts, int *ld, int *device, double **d_slab, int **d_mask, cudaStream_t *stream)
{
    *npts = s->npts; *ld = s->ld; *device = s->device; *d_slab = s->d_slab; *d_mask = s->d_mask; *stream = s->stream;
    return 0;
}
int series_info(const semic_series *s, int *npts, int *ld, int *nvar, int *ndays, double **d)
{
    *npts = s->npts; *ld = s->ld; *nvar = s->nvar; *ndays = s->ndays; *d = s->d;
    return 0;
}
}  // namespace semic

// ------------------------------------------------------------------------------------------------
// small device kernels of this translation unit
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// j-th uniform of the splitmix64 stream keyed by (seed, point, day, var)
__device__ __forceinline__ double uni(uint64_t seed, uint64_t point, uint32_t day, uint32_t var, uint32_t j)
{
    const uint64_t key = seed ^ (((point << 32) + (uint64_t)day) * 16ull + (uint64_t)var);
    const uint64_t x = mix64(key + (uint64_t)(j + 1) * 0x9E3779B97F4A7C15ull);
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double gauss4(uint64_t seed, uint64_t point, uint32_t day, uint32_t var)
{
    const double s = uni(seed, point, day, var, 0) + uni(seed, point, day, var, 1) + uni(seed, point, day, var, 2) +
                     uni(seed, point, day, var, 3);
    return (s - 2.0) * 1.7320508075688772;
}

// SURVEY.md 8(d) synthetic forcing; one thread per point, all days
__global__ void k_synth_forcing(double *f, int npts, int ld, int ndays, uint64_t seed, long long point0, int dayofs,
                                int daystride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const uint64_t pt = (uint64_t)(point0 + i);
    const double tbar = 245.0 + 23.0 * uni(seed, pt, 0xFFFFFFFFu, 15, 0);
    const double sp0 = 63000.0 + 37000.0 * uni(seed, pt, 0xFFFFFFFFu, 14, 0);
    const double t0 = 273.15, eps = 0.62197, pi = 3.141592653589793;
    for (int d = 0; d < ndays; ++d) {
        const uint32_t day = (uint32_t)(dayofs + d * daystride);
        const int doy = (int)(day % 365u);
        const double t2m = tbar + 12.0 * cos(2.0 * pi * (doy - 200) / 365.0) + 3.0 * gauss4(seed, pt, day, SEMIC_V_T2M);
        double swd = 0.0;
        if (doy >= 35 && doy <= 330)
            swd = fmax(0.0, 400.0 * sin(pi * (doy - 35) / 295.0)) * (0.6 + 0.4 * uni(seed, pt, day, SEMIC_V_SWD, 0));
        const double lwd = fmin(340.0, fmax(70.0, 200.0 + 4.0 * (t2m - 260.0) + 15.0 * gauss4(seed, pt, day, SEMIC_V_LWD)));
        const double wind = fmin(22.0, fmax(0.1, 6.5 + 3.0 * gauss4(seed, pt, day, SEMIC_V_WIND)));
        const double sp = sp0 + 1200.0 * gauss4(seed, pt, day, SEMIC_V_SP);
        const double rhoa = sp / (287.05 * t2m);
        const double x = t2m - t0;
        const double es = (t2m < t0) ? 611.2 * exp(22.46 * x / (272.62 + x)) : 611.2 * exp(17.62 * x / (243.12 + x));
        const double qsat = es * eps / (es * (eps - 1.0) + sp);
        const double qq = (0.5 + 0.45 * uni(seed, pt, day, SEMIC_V_QQ, 0)) * qsat;
        double sf = 0.0, rf = 0.0;
        if (t2m < t0 && uni(seed, pt, day, SEMIC_V_SF, 0) >= 0.65)
            sf = -3.5e-8 * log(1.0 - uni(seed, pt, day, SEMIC_V_SF, 1));
        if (t2m > 268.0 && uni(seed, pt, day, SEMIC_V_RF, 0) >= 0.6)
            rf = -6.0e-9 * log(1.0 - uni(seed, pt, day, SEMIC_V_RF, 1));
        // tiled series layout: ((day*ntl + tile)*nvar + var)*32 + lane
        double *row = f + (((size_t)d * (size_t)(ld / kTile) + (size_t)(i / kTile)) * SEMIC_NFORCING) * kTile + (i % kTile);
        row[SEMIC_V_SF * kTile] = sf;
        row[SEMIC_V_RF * kTile] = rf;
        row[SEMIC_V_SWD * kTile] = swd;
        row[SEMIC_V_LWD * kTile] = lwd;
        row[SEMIC_V_WIND * kTile] = wind;
        row[SEMIC_V_SP * kTile] = sp;
        row[SEMIC_V_RHOA * kTile] = rhoa;
        row[SEMIC_V_QQ * kTile] = qq;
        row[SEMIC_V_T2M * kTile] = t2m;
    }
}

__global__ void k_synth_mask(int *mask, int npts, uint64_t seed, long long point0, double frac_ice, double frac_land)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const double u = uni(seed, (uint64_t)(point0 + i), 0xFFFFFFFFu, 13, 0);
    mask[i] = u < frac_ice ? 2 : (u < frac_ice + frac_land ? 1 : 0);
}

__global__ void k_fill(double *p, int n, double v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// SoA rows [nd][nvar][ld_soa] (points [p0, p0+np) of the series) <-> tiled series [..][ld/32][nvar][32]
__global__ void k_soa_to_tiled(const double *soa, long long ld_soa, double *tiled, int ld, int nvar, int day0, int nd, int np,
                               int p0 = 0)
{
    const long long n = (long long)nd * nvar * np;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(k % np);
        const int v = (int)((k / np) % nvar);
        const int d = (int)(k / ((long long)np * nvar));
        const int i = p0 + j;
        tiled[(((long long)(day0 + d) * (ld / kTile) + i / kTile) * nvar + v) * kTile + (i % kTile)] =
            soa[((long long)d * nvar + v) * ld_soa + j];
    }
}
__global__ void k_tiled_to_soa(const double *tiled, int ld, int nvar, int day0, int nd, int p0, int np, double *soa,
                               long long ld_soa)
{
    const long long n = (long long)nd * nvar * np;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(k % np);
        const int v = (int)((k / np) % nvar);
        const int d = (int)(k / ((long long)np * nvar));
        const int i = p0 + j;
        soa[((long long)d * nvar + v) * ld_soa + j] =
            tiled[(((long long)(day0 + d) * (ld / kTile) + i / kTile) * nvar + v) * kTile + (i % kTile)];
    }
}

// the 8 forcing variables sf..qq of one day of a tiled forcing series -> the state slab's forcing fields (t2m, the 9th,
// is written by the step kernel itself: the reference mutates it, surface_physics.f90:209-211)
__constant__ int c_forcing_dst_field[kForcingRows] = {SEMIC_F_SF, SEMIC_F_RF, SEMIC_F_SWD, SEMIC_F_LWD, SEMIC_F_WIND,
                                                      SEMIC_F_SP, SEMIC_F_RHOA, SEMIC_F_QQ, -1};
__global__ void k_tiled_day_to_slab(const double *tiled, int ld_series, int day, int np, double *slab, long long ld_slab)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double *src = tiled + ((long long)day * (ld_series / kTile) + i / kTile) * (kForcingRows * kTile) + (i % kTile);
#pragma unroll
    for (int v = 0; v < kForcingRows; ++v) {
        const int f = c_forcing_dst_field[v];
        if (f >= 0) slab[(long long)f * ld_slab + i] = src[v * kTile];
    }
}

// field_average, surface_physics.f90:601-629, over the 19 averaged fields (:574-593)
__constant__ int c_avg_fields[19] = {SEMIC_F_T2M, SEMIC_F_TSURF, SEMIC_F_HSNOW, SEMIC_F_ALB, SEMIC_F_MELT, SEMIC_F_REFR,
                                     SEMIC_F_SMB, SEMIC_F_ACC, SEMIC_F_LHF, SEMIC_F_SHF, SEMIC_F_LWU, SEMIC_F_SF,
                                     SEMIC_F_RF, SEMIC_F_SP, SEMIC_F_LWD, SEMIC_F_SWD, SEMIC_F_WIND, SEMIC_F_RHOA,
                                     SEMIC_F_QQ};
__global__ void k_average(double *ave, const double *now, int npts, int ld_ave, int ld_now, int step, double nt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
#pragma unroll
    for (int k = 0; k < 19; ++k) {
        const int f = c_avg_fields[k];
        double *a = ave + (size_t)f * ld_ave + i;
        if (step == 0) *a = 0.0;                                          // :611
        else if (step == 1) *a = *a + now[(size_t)f * ld_now + i];        // :614
        else *a = *a / nt;                                                // :621
    }
}

// roofline probes
__global__ void k_probe_dfma(double *out, int iters)
{
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double b = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
__global__ void k_probe_copy(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------
extern "C" int semic_surface_alloc(semic_state **now, int npts)
{
    if (!now || npts < 0) { set_error("surface_alloc: bad arguments"); return SEMIC_ERR_ARG; }
    *now = nullptr;
    semic_state *s = new semic_state();
    s->npts = npts;
    s->ld = round_up(npts > 0 ? npts : 1, kLdAlign);
    cudaError_t e = cudaGetDevice(&s->device);
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "surface_alloc: cudaGetDevice"); }
    const size_t slab_bytes = (size_t)SEMIC_NFIELDS * s->ld * sizeof(double);
    const size_t mask_bytes = (size_t)s->ld * sizeof(int);
    auto fail = [&](cudaError_t err, const char *what) {
        if (s->d_slab) cudaFree(s->d_slab);
        if (s->d_mask) cudaFree(s->d_mask);
        if (s->h_slab) cudaFreeHost(s->h_slab);
        if (s->h_mask) cudaFreeHost(s->h_mask);
        if (s->stream) cudaStreamDestroy(s->stream);
        if (s->ev0) cudaEventDestroy(s->ev0);
        if (s->ev1) cudaEventDestroy(s->ev1);
        delete s;
        return cuda_fail(err, what);
    };
    if ((e = cudaMalloc(&s->d_slab, slab_bytes)) != cudaSuccess) return fail(e, "surface_alloc: cudaMalloc(slab)");
    if ((e = cudaMalloc(&s->d_mask, mask_bytes)) != cudaSuccess) return fail(e, "surface_alloc: cudaMalloc(mask)");
    if ((e = cudaMallocHost(&s->h_slab, slab_bytes)) != cudaSuccess) return fail(e, "surface_alloc: cudaMallocHost(slab)");
    if ((e = cudaMallocHost(&s->h_mask, mask_bytes)) != cudaSuccess) return fail(e, "surface_alloc: cudaMallocHost(mask)");
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "surface_alloc: stream");
    if ((e = cudaEventCreate(&s->ev0)) != cudaSuccess) return fail(e, "surface_alloc: event");
    if ((e = cudaEventCreate(&s->ev1)) != cudaSuccess) return fail(e, "surface_alloc: event");
    // Q8: the reference leaves contents undefined and reads qmr_res before writing it; zero-fill
    std::memset(s->h_slab, 0, slab_bytes);
    std::memset(s->h_mask, 0, mask_bytes);
    if ((e = cudaMemsetAsync(s->d_slab, 0, slab_bytes, s->stream)) != cudaSuccess) return fail(e, "surface_alloc: memset");
    if ((e = cudaMemsetAsync(s->d_mask, 0, mask_bytes, s->stream)) != cudaSuccess) return fail(e, "surface_alloc: memset");
    if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess) return fail(e, "surface_alloc: sync");
    *now = s;
    return SEMIC_OK;
}

extern "C" int semic_surface_dealloc(semic_state *s)
{
    if (!s) { set_error("surface_dealloc: null state"); return SEMIC_ERR_ARG; }
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    cudaFree(s->d_slab);
    cudaFree(s->d_mask);
    cudaFreeHost(s->h_slab);
    cudaFreeHost(s->h_mask);
    cudaEventDestroy(s->ev0);
    cudaEventDestroy(s->ev1);
    for (cudaEvent_t e : s->ev_up) cudaEventDestroy(e);
    for (cudaEvent_t e : s->ev_done) cudaEventDestroy(e);
    if (s->s_in) cudaStreamDestroy(s->s_in);
    if (s->s_out) cudaStreamDestroy(s->s_out);
    cudaStreamDestroy(s->stream);
    delete s;
    return SEMIC_OK;
}

extern "C" int semic_surface_physics_par_load(semic_par *par, const char *filename)
{
    if (!par || !filename) { set_error("surface_physics_par_load: bad arguments"); return SEMIC_ERR_ARG; }
    return namelist_load(par, filename);
}

extern "C" int semic_surface_boundary_define(semic_bnd *bnd, const char *boundary, int n, int len)
{
    if (!bnd || (n > 0 && !boundary) || len <= 0) { set_error("surface_boundary_define: bad arguments"); return SEMIC_ERR_ARG; }
    std::memset(bnd, 0, sizeof *bnd);                                       // :829-840
    for (int q = 0; q < n; ++q) {                                           // :843-874
        const std::string name = ftrim(boundary + (size_t)q * len, len);
        if (name == "t2m") bnd->t2m = 1;
        else if (name == "tsurf") bnd->tsurf = 1;
        else if (name == "hsnow") bnd->hsnow = 1;
        else if (name == "alb") bnd->alb = 1;
        else if (name == "melt") bnd->melt = 1;
        else if (name == "refr") bnd->refr = 1;
        else if (name == "smb") bnd->smb = 1;
        else if (name == "acc") bnd->acc = 1;
        else if (name == "lhf") bnd->lhf = 1;
        else if (name == "shf") bnd->shf = 1;
        else if (name == "subl") bnd->subl = 1;
        else if (name == "amp") bnd->amp = 1;
        // default: pass (:871-872)
    }
    return SEMIC_OK;
}

// Fortran G13.6 edit descriptor (print_param, surface_physics.f90:896-913)
static void g13_6(double x, char *out)
{
    const double ax = std::fabs(x);
    if (ax != 0.0 && (ax < 0.1 - 0.5e-7 || ax >= 1.0e6 - 0.5)) {
        int e = (int)std::floor(std::log10(ax)) + 1;
        double m = x / std::pow(10.0, e);
        if (std::fabs(m) >= 0.9999995) { m /= 10.0; ++e; }
        char tmp[64];
        snprintf(tmp, sizeof tmp, "%.6fE%c%02d", m, e < 0 ? '-' : '+', std::abs(e));
        snprintf(out, 32, "%13.31s", tmp);
    } else {
        int k = ax == 0.0 ? 1 : (int)std::floor(std::log10(ax)) + 1;      // digits before the point
        if (ax != 0.0 && ax >= std::pow(10.0, k) - 0.5 * std::pow(10.0, k - 6)) ++k;
        int dec = 6 - k;
        if (dec < 0) dec = 0;
        char tmp[64];
        snprintf(tmp, sizeof tmp, "%.*f", dec, x);
        if (dec == 0) strncat(tmp, ".", sizeof tmp - strlen(tmp) - 1);
        snprintf(out, 32, "%9s    ", tmp);
    }
}

extern "C" int semic_print_param(const semic_par *par)
{
    if (!par) { set_error("print_param: null"); return SEMIC_ERR_ARG; }
    for (int q = 0; q < SEMIC_NBOUNDARY; ++q) {                             // :887-892
        const std::string b = ftrim(par->boundary[q], SEMIC_STRLEN);
        if (!b.empty() && b.size() != (size_t)SEMIC_STRLEN) printf("boundary %s\n", b.c_str());
    }
    printf("alb_scheme  %s\n", ftrim(par->alb_scheme, SEMIC_STRLEN).c_str());
    printf("nx         %5d\n", par->nx);
    printf("n_ksub     %5d\n", par->n_ksub);
    const struct { const char *n; double v; } rows[] = {
        {"ceff       ", par->ceff}, {"albi       ", par->albi}, {"albl       ", par->albl},
        {"alb_smax   ", par->alb_smax}, {"alb_smin   ", par->alb_smin}, {"hcrit      ", par->hcrit},
        {"rcrit      ", par->rcrit}, {"amp        ", par->amp}, {"csh        ", par->csh},
        {"tmin       ", par->tmin}, {"tmax       ", par->tmax}, {"tstic      ", par->tstic},
        {"tsticsub   ", par->tsticsub}, {"clh        ", par->clh}, {"tau_a      ", par->tau_a},
        {"tau_f      ", par->tau_f}, {"w_crit     ", par->w_crit}, {"mcrit      ", par->mcrit}};
    for (const auto &r : rows) {
        char g[32];
        g13_6(r.v, g);
        printf("%s%s\n", r.n, g);
    }
    fflush(stdout);
    return SEMIC_OK;
}

extern "C" int semic_print_boundary_opt(const semic_bnd *b)
{
    if (!b) { set_error("print_boundary_opt: null"); return SEMIC_ERR_ARG; }
    const struct { const char *n; int v; } rows[] = {                       // :921-931
        {"tsurf   ", b->tsurf}, {"hsnow   ", b->hsnow}, {"alb     ", b->alb}, {"melt    ", b->melt},
        {"refr    ", b->refr}, {"acc     ", b->acc}, {"lhf     ", b->lhf}, {"shf     ", b->shf},
        {"subl    ", b->subl}, {"smb     ", b->smb}, {"amp     ", b->amp}};
    for (const auto &r : rows) printf("%s%c\n", r.n, r.v ? 'T' : 'F');
    fflush(stdout);
    return SEMIC_OK;
}

extern "C" double semic_constant(const char *name)
{
    const std::string n = name ? name : "";
    if (n == "sigm") return 5.67e-8;                                        // :16
    if (n == "cap") return 1000.0;                                          // :22
    if (n == "eps") return 0.62197;                                         // :17
    if (n == "cls") return 2.83e6;                                          // :19
    if (n == "clv") return 2.5e6;                                           // :21
    if (n == "clm") return 3.30e5;
    if (n == "rhow") return 1000.0;
    if (n == "t0") return 273.15;
    if (n == "hsmax") return 5.0;
    if (n == "pi") return 3.141592653589793238462643;
    return std::nan("");
}

// ------------------------------------------------------------------------------------------------
// state access
// ------------------------------------------------------------------------------------------------
extern "C" int semic_state_npts(const semic_state *s) { return s ? s->npts : SEMIC_ERR_ARG; }
extern "C" int semic_state_ld(const semic_state *s) { return s ? s->ld : SEMIC_ERR_ARG; }
extern "C" double *semic_state_field(semic_state *s, int field)
{
    if (!s || field < 0 || field >= SEMIC_NFIELDS) { set_error("state_field: bad arguments"); return nullptr; }
    return s->h_slab + (size_t)field * s->ld;
}
extern "C" int *semic_state_mask(semic_state *s) { return s ? s->h_mask : nullptr; }
extern "C" double *semic_state_device_slab(semic_state *s) { return s ? s->d_slab : nullptr; }

static int state_copy(semic_state *s, uint64_t fields, bool up, int p0 = 0, int np = -1, cudaStream_t st = nullptr)
{
    CU(cudaSetDevice(s->device));
    if (np < 0) np = s->npts;
    if (!st) st = s->stream;
    const size_t rowb = (size_t)np * sizeof(double);
    int f = 0;
    while (f < SEMIC_NFIELDS) {                       // coalesce runs of adjacent fields into one copy
        if (!((fields >> f) & 1ull)) { ++f; continue; }
        int g = f;
        while (g + 1 < SEMIC_NFIELDS && ((fields >> (g + 1)) & 1ull)) ++g;
        const size_t off = (size_t)f * s->ld + p0;
        const size_t nrows = (size_t)(g - f + 1);
        if (up) CU(cudaMemcpy2DAsync(s->d_slab + off, (size_t)s->ld * 8, s->h_slab + off, (size_t)s->ld * 8, rowb, nrows,
                                     cudaMemcpyHostToDevice, st));
        else    CU(cudaMemcpy2DAsync(s->h_slab + off, (size_t)s->ld * 8, s->d_slab + off, (size_t)s->ld * 8, rowb, nrows,
                                     cudaMemcpyDeviceToHost, st));
        f = g + 1;
    }
    if ((fields >> SEMIC_F_MASK) & 1ull) {
        if (up) CU(cudaMemcpyAsync(s->d_mask + p0, s->h_mask + p0, (size_t)np * sizeof(int), cudaMemcpyHostToDevice, st));
        else    CU(cudaMemcpyAsync(s->h_mask + p0, s->d_mask + p0, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    return SEMIC_OK;
}

extern "C" int semic_state_upload(semic_state *s, uint64_t fields)
{
    if (!s) { set_error("state_upload: null state"); return SEMIC_ERR_ARG; }
    int rc = state_copy(s, fields, true);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return SEMIC_OK;
}
extern "C" int semic_state_download(semic_state *s, uint64_t fields)
{
    if (!s) { set_error("state_download: null state"); return SEMIC_ERR_ARG; }
    int rc = state_copy(s, fields, false);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return SEMIC_OK;
}
extern "C" int semic_state_set_sync(semic_state *s, uint64_t up, uint64_t down)
{
    if (!s) { set_error("state_set_sync: null state"); return SEMIC_ERR_ARG; }
    s->up_fields = up;
    s->down_fields = down;
    return SEMIC_OK;
}

extern "C" int semic_state_fill(semic_state *s, int field, double value)
{
    if (!s || field < 0 || field >= SEMIC_NFIELDS) { set_error("state_fill: bad arguments"); return SEMIC_ERR_ARG; }
    CU(cudaSetDevice(s->device));
    double *h = s->h_slab + (size_t)field * s->ld;
    for (int i = 0; i < s->npts; ++i) h[i] = value;
    if (s->npts > 0) {
        k_fill<<<(s->npts + 255) / 256, 256, 0, s->stream>>>(s->d_slab + (size_t)field * s->ld, s->npts, value);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s->stream));
    }
    return SEMIC_OK;
}

extern "C" int semic_state_synth_mask(semic_state *s, uint64_t seed, int64_t point0, double frac_ice, double frac_land)
{
    if (!s) { set_error("state_synth_mask: null state"); return SEMIC_ERR_ARG; }
    CU(cudaSetDevice(s->device));
    if (s->npts > 0) {
        k_synth_mask<<<(s->npts + 255) / 256, 256, 0, s->stream>>>(s->d_mask, s->npts, seed, (long long)point0, frac_ice, frac_land);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(s->h_mask, s->d_mask, (size_t)s->npts * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return SEMIC_OK;
}

// ------------------------------------------------------------------------------------------------
// series
// ------------------------------------------------------------------------------------------------
extern "C" int semic_series_alloc(semic_series **out, int npts, int nvar, int ndays)
{
    if (!out || npts < 0 || nvar <= 0 || ndays <= 0) { set_error("series_alloc: bad arguments"); return SEMIC_ERR_ARG; }
    *out = nullptr;
    semic_series *s = new semic_series();
    s->npts = npts; s->nvar = nvar; s->ndays = ndays;
    s->ld = round_up(npts > 0 ? npts : 1, kLdAlign);
    cudaError_t e = cudaGetDevice(&s->device);
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "series_alloc: cudaGetDevice"); }
    const size_t bytes = (size_t)ndays * nvar * s->ld * sizeof(double);
    if ((e = cudaMalloc(&s->d, bytes)) != cudaSuccess) { delete s; return cuda_fail(e, "series_alloc: cudaMalloc"); }
    if ((e = cudaMemset(s->d, 0, bytes)) != cudaSuccess) { cudaFree(s->d); delete s; return cuda_fail(e, "series_alloc: memset"); }
    // the fill runs on the legacy stream; the kernels that read the series run on non-blocking streams, which are not
    // ordered behind it: finish it here
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) { cudaFree(s->d); delete s; return cuda_fail(e, "series_alloc: sync"); }
    *out = s;
    return SEMIC_OK;
}
extern "C" int semic_series_free(semic_series *s)
{
    if (!s) { set_error("series_free: null"); return SEMIC_ERR_ARG; }
    cudaSetDevice(s->device);
    cudaFree(s->d);
    delete s;
    return SEMIC_OK;
}
extern "C" int semic_series_npts(const semic_series *s) { return s ? s->npts : SEMIC_ERR_ARG; }
extern "C" int semic_series_ld(const semic_series *s) { return s ? s->ld : SEMIC_ERR_ARG; }
extern "C" int semic_series_nvar(const semic_series *s) { return s ? s->nvar : SEMIC_ERR_ARG; }
extern "C" int semic_series_ndays(const semic_series *s) { return s ? s->ndays : SEMIC_ERR_ARG; }
extern "C" double *semic_series_device_ptr(semic_series *s) { return s ? s->d : nullptr; }

// Host arrays are SoA per day ([ndays][nvar][host_ld], the layout of the reference's files and arrays); the device
// series is tile-interleaved, so transfers go through SoA staging buffers and a transpose kernel.  Units of
// (one day) x (a slice of points) are double-buffered over a copy stream and a transpose stream, so that the PCIe copy of
// one unit overlaps the transpose of the previous one; the staging buffers are kept per device between calls.
namespace {
struct Staging {
    double *buf[2] = {nullptr, nullptr};
    size_t cap = 0;                                  // doubles per buffer
    cudaStream_t copy = nullptr, conv = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_conv[2] = {nullptr, nullptr};
};
std::mutex g_stage_mu[64];                          // one per device: shards on different GPUs transfer concurrently
Staging g_stage[64];
constexpr int kStageSlice = 1 << 21;                 // points per unit (2 Mi): 151 MB of forcing, 134 MB of output per buffer

cudaError_t staging_get(int device, size_t need, Staging **out)
{
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    Staging &g = g_stage[device];
    cudaError_t e;
    if (!g.copy) {
        if ((e = cudaStreamCreateWithFlags(&g.copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&g.conv, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int b = 0; b < 2; ++b) {
            if ((e = cudaEventCreateWithFlags(&g.ev_copy[b], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&g.ev_conv[b], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
    }
    if (g.cap < need) {
        for (int b = 0; b < 2; ++b) {
            if (g.buf[b]) cudaFree(g.buf[b]);
            g.buf[b] = nullptr;
            if ((e = cudaMalloc(&g.buf[b], need * sizeof(double))) != cudaSuccess) { g.cap = 0; return e; }
        }
        g.cap = need;
    }
    *out = &g;
    return cudaSuccess;
}
}  // namespace

extern "C" int semic_series_upload(semic_series *s, int day0, int ndays, const double *host, int host_ld)
{
    if (!s || !host || day0 < 0 || ndays < 0 || day0 + ndays > s->ndays || host_ld < s->npts) {
        set_error("series_upload: bad arguments");
        return SEMIC_ERR_ARG;
    }
    CU(cudaSetDevice(s->device));
    if (ndays == 0 || s->npts == 0) return SEMIC_OK;
    if (s->device < 0 || s->device >= 64) { set_error("series_upload: bad device"); return SEMIC_ERR_ARG; }
    std::lock_guard<std::mutex> lock(g_stage_mu[s->device]);
    const int slice = s->npts < kStageSlice ? s->npts : kStageSlice;
    Staging *g = nullptr;
    CU(staging_get(s->device, (size_t)s->nvar * slice, &g));
    CU(cudaDeviceSynchronize());                      // whoever wrote or reads the series before this call is done
    long long unit = 0;
    cudaError_t e = cudaSuccess;
    for (int d = 0; d < ndays && e == cudaSuccess; ++d) {
        for (int p0 = 0; p0 < s->npts && e == cudaSuccess; p0 += slice, ++unit) {
            const int b = (int)(unit & 1);
            const int np = s->npts - p0 < slice ? s->npts - p0 : slice;
            e = cudaStreamWaitEvent(g->copy, g->ev_conv[b], 0);                     // the transpose that last read buf[b]
            if (e == cudaSuccess)
                e = cudaMemcpy2DAsync(g->buf[b], (size_t)np * 8, host + (size_t)d * s->nvar * host_ld + p0, (size_t)host_ld * 8,
                                      (size_t)np * 8, (size_t)s->nvar, cudaMemcpyHostToDevice, g->copy);
            if (e == cudaSuccess) e = cudaEventRecord(g->ev_copy[b], g->copy);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(g->conv, g->ev_copy[b], 0);
            if (e == cudaSuccess) {
                k_soa_to_tiled<<<1184, 256, 0, g->conv>>>(g->buf[b], np, s->d, s->ld, s->nvar, day0 + d, 1, np, p0);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaEventRecord(g->ev_conv[b], g->conv);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->conv);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->copy);
    if (e != cudaSuccess) return cuda_fail(e, "series_upload");
    return SEMIC_OK;
}
extern "C" int semic_series_download(const semic_series *s, int day0, int ndays, double *host, int host_ld, int point0,
                                     int npoints)
{
    if (!s || !host || day0 < 0 || ndays < 0 || day0 + ndays > s->ndays || point0 < 0 || npoints < 0 ||
        point0 + npoints > s->npts || host_ld < npoints) {
        set_error("series_download: bad arguments");
        return SEMIC_ERR_ARG;
    }
    CU(cudaSetDevice(s->device));
    if (ndays == 0 || npoints == 0) return SEMIC_OK;
    if (s->device < 0 || s->device >= 64) { set_error("series_download: bad device"); return SEMIC_ERR_ARG; }
    std::lock_guard<std::mutex> lock(g_stage_mu[s->device]);
    const int slice = npoints < kStageSlice ? npoints : kStageSlice;
    Staging *g = nullptr;
    CU(staging_get(s->device, (size_t)s->nvar * slice, &g));
    CU(cudaDeviceSynchronize());                      // the kernels that wrote the series are done
    long long unit = 0;
    cudaError_t e = cudaSuccess;
    for (int d = 0; d < ndays && e == cudaSuccess; ++d) {
        for (int q0 = 0; q0 < npoints && e == cudaSuccess; q0 += slice, ++unit) {
            const int b = (int)(unit & 1);
            const int np = npoints - q0 < slice ? npoints - q0 : slice;
            e = cudaStreamWaitEvent(g->conv, g->ev_copy[b], 0);                     // the copy that last read buf[b]
            if (e == cudaSuccess) {
                k_tiled_to_soa<<<1184, 256, 0, g->conv>>>(s->d, s->ld, s->nvar, day0 + d, 1, point0 + q0, np, g->buf[b], np);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaEventRecord(g->ev_conv[b], g->conv);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(g->copy, g->ev_conv[b], 0);
            if (e == cudaSuccess)
                e = cudaMemcpy2DAsync(host + (size_t)d * s->nvar * host_ld + q0, (size_t)host_ld * 8, g->buf[b], (size_t)np * 8,
                                      (size_t)np * 8, (size_t)s->nvar, cudaMemcpyDeviceToHost, g->copy);
            if (e == cudaSuccess) e = cudaEventRecord(g->ev_copy[b], g->copy);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->copy);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->conv);
    if (e != cudaSuccess) return cuda_fail(e, "series_download");
    return SEMIC_OK;
}

extern "C" int semic_series_synth_forcing(semic_series *s, uint64_t seed, int64_t point0, int dayofs, int daystride)
{
    if (!s || s->nvar != SEMIC_NFORCING) { set_error("series_synth_forcing: series must have 9 variables"); return SEMIC_ERR_ARG; }
    CU(cudaSetDevice(s->device));
    if (s->npts > 0) {
        k_synth_forcing<<<(s->npts + 127) / 128, 128>>>(s->d, s->npts, s->ld, s->ndays, seed, (long long)point0, dayofs,
                                                        daystride > 0 ? daystride : 1);
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
    }
    return SEMIC_OK;
}

// ------------------------------------------------------------------------------------------------
// the time step
// ------------------------------------------------------------------------------------------------
static int launches_total = 0;

static int check_par(const semic_par *par, const char *who)
{
    if (!par) { set_error("%s: null parameters", who); return SEMIC_ERR_ARG; }
    if (par->n_ksub < 0) { set_error("%s: n_ksub < 0", who); return SEMIC_ERR_ARG; }
    return SEMIC_OK;
}

// one fused launch on the state's own slab (forcing = the state's SoA forcing fields)
static int step_on_slab(semic_state *s, const semic_par *par, const semic_bnd *bnd, int eb_steps, int do_mb, bool whole_step,
                        int strict, int p0 = 0, int np = -1)
{
    RunArgs a;
    std::memset(&a, 0, sizeof a);
    if (np < 0) np = s->npts;
    a.slab = s->d_slab + p0; a.mask = s->d_mask + p0; a.nx = np; a.ld = s->ld; a.sld = s->ld;
    a.forcing = s->d_slab + p0;
    a.f_tiled = 0; a.f_nwin = 1; a.f_day0 = 0;
    const int src_field[kForcingRows] = {SEMIC_F_SF, SEMIC_F_RF, SEMIC_F_SWD, SEMIC_F_LWD, SEMIC_F_WIND, SEMIC_F_SP,
                                         SEMIC_F_RHOA, SEMIC_F_QQ, SEMIC_F_T2M};
    for (int v = 0; v < kForcingRows; ++v) a.f_var_off[v] = (long long)src_field[v] * s->ld;
    a.vali = nullptr; a.out = nullptr; a.o_nring = 1; a.o_nvar = SEMIC_NOUT; a.out_start = 0;
    a.ndays = 1;
    a.eb_steps = eb_steps; a.do_mb = do_mb;
    a.P = make_devpar(par);
    a.B = *bnd;
    const bool generic = !whole_step || any_consulted_flag(bnd) || par->n_ksub < 1;
    cudaError_t e = strict ? launch_run_strict(a, generic, s->stream, &launches_total)
                           : launch_run_fast(a, generic, s->stream, &launches_total);
    if (e != cudaSuccess) return cuda_fail(e, "time step kernel launch");
    return SEMIC_OK;
}

static int dropin(semic_state *s, const semic_par *par, const semic_bnd *bnd, int eb_steps, int do_mb, bool whole, const char *who)
{
    int rc = check_par(par, who);
    if (rc) return rc;
    if (!s || !bnd) { set_error("%s: null argument", who); return SEMIC_ERR_ARG; }
    CU(cudaSetDevice(s->device));
    if (s->npts == 0) return SEMIC_OK;
    const int strict = getenv("SEMIC_STRICT") ? 1 : 0;
    // Large grids: the points are independent, so the call is pipelined over slices of the grid -- upload of slice
    // i+1, kernel of slice i and download of slice i-1 run concurrently (PCIe is full duplex): the call costs about
    // max(bytes up, bytes down) / PCIe bandwidth instead of their sum.
    int slice = 1 << 19;
    if (const char *e = getenv("SEMIC_STEP_SLICE")) { const int v = atoi(e); if (v > 0) slice = round_up(v, kLdAlign); }
    const int nslices = (s->npts + slice - 1) / slice;
    if (nslices < 2) {
        if ((rc = state_copy(s, s->up_fields, true))) return rc;
        if ((rc = step_on_slab(s, par, bnd, eb_steps, do_mb, whole, strict))) return rc;
        if ((rc = state_copy(s, s->down_fields, false))) return rc;
        CU(cudaStreamSynchronize(s->stream));
        return SEMIC_OK;
    }
    if (!s->s_in) CU(cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking));
    if (!s->s_out) CU(cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking));
    while ((int)s->ev_up.size() < nslices) {
        cudaEvent_t a = nullptr, b = nullptr;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        s->ev_up.push_back(a);
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        s->ev_done.push_back(b);
    }
    CU(cudaEventRecord(s->ev0, s->stream));                       // work queued on the state's stream so far comes first
    CU(cudaStreamWaitEvent(s->s_in, s->ev0, 0));
    for (int i = 0; i < nslices; ++i) {
        const int p0 = i * slice;
        const int np = (p0 + slice <= s->npts) ? slice : s->npts - p0;
        if ((rc = state_copy(s, s->up_fields, true, p0, np, s->s_in))) return rc;
        CU(cudaEventRecord(s->ev_up[i], s->s_in));
        CU(cudaStreamWaitEvent(s->stream, s->ev_up[i], 0));
        if ((rc = step_on_slab(s, par, bnd, eb_steps, do_mb, whole, strict, p0, np))) return rc;
        CU(cudaEventRecord(s->ev_done[i], s->stream));
        CU(cudaStreamWaitEvent(s->s_out, s->ev_done[i], 0));
        if ((rc = state_copy(s, s->down_fields, false, p0, np, s->s_out))) return rc;
    }
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaStreamSynchronize(s->s_out));
    return SEMIC_OK;
}

extern "C" int semic_surface_energy_and_mass_balance(semic_state *now, const semic_par *par, const semic_bnd *bnd, int day,
                                                     int year)
{
    (void)day; (void)year;                                                  // dead arguments in the reference too (:143-144)
    return dropin(now, par, bnd, par ? par->n_ksub : 0, 1, true, "surface_energy_and_mass_balance");
}
extern "C" int semic_energy_balance(semic_state *now, const semic_par *par, const semic_bnd *bnd, int day, int year)
{
    (void)day; (void)year;
    return dropin(now, par, bnd, 1, 0, false, "energy_balance");
}
extern "C" int semic_mass_balance(semic_state *now, const semic_par *par, const semic_bnd *bnd, int day, int year)
{
    (void)day; (void)year;
    return dropin(now, par, bnd, 0, 1, false, "mass_balance");
}

// forcing/vali/out are TILED device series with the state's leading dimension
static void fill_run_args(RunArgs &a, semic_state *now, const semic_par *par, const semic_bnd *bnd, const double *forcing,
                          int f_nwin, const double *vali, int day0, int ndays, double *out, int o_nring, int o_nvar,
                          int out_days)
{
    std::memset(&a, 0, sizeof a);
    a.slab = now->d_slab; a.mask = now->d_mask; a.nx = now->npts; a.ld = now->ld; a.sld = now->ld;
    a.forcing = forcing;
    a.f_tiled = 1;
    a.f_nwin = f_nwin;
    a.f_day0 = ((day0 % f_nwin) + f_nwin) % f_nwin;
    a.vali = vali;
    a.out = out_days > 0 ? out : nullptr;
    a.o_nring = o_nring > 0 ? o_nring : 1;
    a.o_nvar = o_nvar > 0 ? o_nvar : SEMIC_NOUT;
    a.out_start = ndays - out_days;
    a.ndays = ndays;
    a.eb_steps = par->n_ksub;
    a.do_mb = 1;
    a.P = make_devpar(par);
    a.B = *bnd;
}

// the state's forcing fields hold the last day's forcing after a multi-day run (as after the reference's day loop)
static int copy_last_forcing(semic_state *now, const double *forcing, int f_nwin, int day0, int ndays)
{
    const int fd = (((day0 + ndays - 1) % f_nwin) + f_nwin) % f_nwin;
    k_tiled_day_to_slab<<<(now->npts + 255) / 256, 256, 0, now->stream>>>(forcing, now->ld, fd, now->npts, now->d_slab, now->ld);
    CU(cudaGetLastError());
    return SEMIC_OK;
}

extern "C" int semic_run(semic_state *now, const semic_par *par, const semic_bnd *bnd, const semic_series *forcing,
                         const semic_series *vali, int day0, int ndays, semic_series *out, const semic_run_opts *opts,
                         float *kernel_ms)
{
    int rc = check_par(par, "semic_run");
    if (rc) return rc;
    if (!now || !bnd || !forcing || ndays < 0) { set_error("semic_run: bad arguments"); return SEMIC_ERR_ARG; }
    if (forcing->nvar != SEMIC_NFORCING || forcing->npts != now->npts) { set_error("semic_run: forcing series must be 9 x npts"); return SEMIC_ERR_ARG; }
    if (vali && (vali->nvar != SEMIC_NOUT || vali->npts != now->npts || vali->ndays != forcing->ndays)) {
        set_error("semic_run: validation series must be 8 x npts with the forcing's window");
        return SEMIC_ERR_ARG;
    }
    const int out_days = (opts && out) ? opts->out_days : 0;
    if (out_days < 0 || out_days > ndays) { set_error("semic_run: out_days out of range"); return SEMIC_ERR_ARG; }
    if (out && (out->npts != now->npts || (out->nvar != SEMIC_NOUT && out->nvar != SEMIC_NOUT + 1))) {
        set_error("semic_run: output series must be 8 (or 9) x npts");
        return SEMIC_ERR_ARG;
    }
    if (forcing->ld != now->ld || (vali && vali->ld != now->ld) || (out && out->ld != now->ld)) {
        set_error("semic_run: leading dimensions differ");
        return SEMIC_ERR_ARG;
    }
    if (kernel_ms) *kernel_ms = 0.f;
    CU(cudaSetDevice(now->device));
    if (now->npts == 0 || ndays == 0) return SEMIC_OK;
    RunArgs a;
    fill_run_args(a, now, par, bnd, forcing->d, forcing->ndays, vali ? vali->d : nullptr, day0, ndays, out ? out->d : nullptr,
                  out ? out->ndays : 1, out ? out->nvar : SEMIC_NOUT, out_days);
    const bool generic = any_consulted_flag(bnd) || par->n_ksub < 1;
    const int strict = opts ? opts->strict : 0;
    if (!strict) CU(prepare_fast());
    a.avg_days = (opts && out_days > 0 && opts->avg_days > 1) ? opts->avg_days : 0;
    if (a.avg_days > 1) {
        if (generic || strict || out->nvar != SEMIC_NOUT) {
            set_error("semic_run: avg_days needs a plain run (no boundary overrides, fast arithmetic, 8-row output)");
            return SEMIC_ERR_ARG;
        }
        if (out->ndays < 1) { set_error("semic_run: empty output series"); return SEMIC_ERR_ARG; }
    }
    CU(cudaEventRecord(now->ev0, now->stream));
    cudaError_t e = strict ? launch_run_strict(a, generic, now->stream, &launches_total)
                           : launch_run_fast(a, generic, now->stream, &launches_total);
    if (e != cudaSuccess) return cuda_fail(e, "semic_run: kernel launch");
    CU(cudaEventRecord(now->ev1, now->stream));
    if ((rc = copy_last_forcing(now, forcing->d, forcing->ndays, day0, ndays))) return rc;
    CU(cudaStreamSynchronize(now->stream));
    if (kernel_ms) CU(cudaEventElapsedTime(kernel_ms, now->ev0, now->ev1));
    return SEMIC_OK;
}

extern "C" void *semic_host_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) { cuda_fail(e, "semic_host_alloc"); return nullptr; }
    return p;
}
extern "C" int semic_host_free(void *p)
{
    CU(cudaFreeHost(p));
    return SEMIC_OK;
}

// Host-streamed run.  The unit of the pipeline is (chunk of days) x (slice of points): H2D of unit u+1 and D2H of unit
// u-1 overlap the kernel of unit u.  Slicing the points (independent, surface_physics.f90:138-374 has no neighbour
// terms) keeps the pipeline's fill and drain -- the first H2D and the last D2H, which nothing can overlap -- at a
// fraction of a day's transfer.  Host buffers are SoA per day (the reference's array layout); each unit is re-tiled
// on the device (2 x 72 B/point-day of HBM traffic, negligible next to PCIe).
extern "C" int semic_run_host(semic_state *now, const semic_par *par, const semic_bnd *bnd, const double *h_forcing,
                              const double *h_vali, int nwin, int ldh, int ndays, double *h_out, int out_days,
                              int chunk_days, const semic_run_opts *opts, float *total_ms)
{
    int rc = check_par(par, "semic_run_host");
    if (rc) return rc;
    if (!now || !bnd || !h_forcing || nwin <= 0 || ldh < (now ? now->npts : 0) || ndays < 0 || out_days < 0 ||
        out_days > ndays || (out_days > 0 && !h_out)) {
        set_error("semic_run_host: bad arguments");
        return SEMIC_ERR_ARG;
    }
    if (total_ms) *total_ms = 0.f;
    CU(cudaSetDevice(now->device));
    if (now->npts == 0 || ndays == 0) return SEMIC_OK;
    if (chunk_days <= 0) chunk_days = 8;
    if (chunk_days > ndays) chunk_days = ndays;
    const int ld = now->ld, npts = now->npts;
    const bool generic = any_consulted_flag(bnd) || par->n_ksub < 1;
    const int strict = opts ? opts->strict : 0;
    // period means instead of daily rows (surface_physics_average protocol, in-kernel): h_out holds ceil(out_days/avg) rows.
    // Periods must not straddle pipeline units: whole periods per chunk, output starting on a chunk boundary.
    const int avg = (opts && opts->avg_days > 1 && out_days > 0) ? opts->avg_days : 1;
    if (avg > 1 && (generic || strict || chunk_days % avg != 0 || (ndays - out_days) % chunk_days != 0)) {
        set_error("semic_run_host: avg_days needs a plain fast-mode run, chunk_days %% avg_days == 0 and (ndays - out_days) %% chunk_days == 0");
        return SEMIC_ERR_ARG;
    }
    const bool use_vali = h_vali != nullptr && generic;
    // point slices: multiples of kLdAlign points, about 4 Mi points each (SEMIC_HOST_SLICE overrides, for tests)
    int slice = 1 << 22;
    if (const char *e = getenv("SEMIC_HOST_SLICE")) { const int v = atoi(e); if (v > 0) slice = round_up(v, kLdAlign); }
    if (slice > ld) slice = ld;
    const int nslices = (npts + slice - 1) / slice;
    const int sld = slice;                                        // leading dimension of the units' tiled buffers

    // per buffer set b: SoA staging (s) and tiled (t) copies of forcing f, vali v, output o
    double *d_fs[2] = {nullptr, nullptr}, *d_ft[2] = {nullptr, nullptr}, *d_vs[2] = {nullptr, nullptr}, *d_vt[2] = {nullptr, nullptr};
    double *d_os[2] = {nullptr, nullptr}, *d_ot[2] = {nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_t[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_o[2] = {nullptr, nullptr};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    const size_t fsb = (size_t)chunk_days * SEMIC_NFORCING * slice * 8, ftb = (size_t)chunk_days * SEMIC_NFORCING * sld * 8;
    const size_t osb = (size_t)chunk_days * SEMIC_NOUT * slice * 8, otb = (size_t)chunk_days * SEMIC_NOUT * sld * 8;
    int status = SEMIC_OK;
    auto cleanup = [&]() {
        for (int b = 0; b < 2; ++b) {
            double *ptrs[6] = {d_fs[b], d_ft[b], d_vs[b], d_vt[b], d_os[b], d_ot[b]};
            for (double *p : ptrs) if (p) cudaFree(p);
            cudaEvent_t evs[4] = {ev_in[b], ev_t[b], ev_k[b], ev_o[b]};
            for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
    };
#define CUX(call)                                                                  \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) { status = cuda_fail(e__, #call); cleanup(); return status; } \
    } while (0)
    for (int b = 0; b < 2; ++b) {
        CUX(cudaMalloc(&d_fs[b], fsb));
        CUX(cudaMalloc(&d_ft[b], ftb));
        CUX(cudaMemset(d_ft[b], 0, ftb));
        if (use_vali) { CUX(cudaMalloc(&d_vs[b], osb)); CUX(cudaMalloc(&d_vt[b], otb)); CUX(cudaMemset(d_vt[b], 0, otb)); }
        if (out_days > 0) { CUX(cudaMalloc(&d_os[b], osb)); CUX(cudaMalloc(&d_ot[b], otb)); }
        CUX(cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&ev_t[b], cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&ev_k[b], cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&ev_o[b], cudaEventDisableTiming));
    }
    CUX(cudaDeviceSynchronize());                 // the legacy-stream memsets above are not ordered before the non-blocking streams
    CUX(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CUX(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    CUX(cudaEventCreate(&t0));
    CUX(cudaEventCreate(&t1));

    if (!strict) CUX(prepare_fast());
    const int nchunks = (ndays + chunk_days - 1) / chunk_days;
    const int out_first = ndays - out_days;
    CUX(cudaEventRecord(t0, now->stream));
    CUX(cudaStreamWaitEvent(s_in, t0, 0));
    long long unit = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int d0 = c * chunk_days;
        const int nd = (d0 + chunk_days <= ndays) ? chunk_days : ndays - d0;
        const int o_lo = out_first > d0 ? out_first : d0;                 // output days of this chunk: [o_lo, d0+nd)
        const int o_n = (d0 + nd > o_lo) ? d0 + nd - o_lo : 0;
        for (int sl = 0; sl < nslices; ++sl, ++unit) {
            const int b = (int)(unit & 1);
            const int p0 = sl * slice;
            const int np = (p0 + slice <= npts) ? slice : npts - p0;
            // H2D of this unit's SoA rows (day d -> window row d % nwin), once the re-tiling that last read the staging is done
            if (unit >= 2) CUX(cudaStreamWaitEvent(s_in, ev_t[b], 0));
            int done = 0;
            while (done < nd) {
                const int w = (d0 + done) % nwin;
                const int run = (nd - done < nwin - w) ? nd - done : nwin - w;
                CUX(cudaMemcpy2DAsync(d_fs[b] + (size_t)done * SEMIC_NFORCING * np, (size_t)np * 8,
                                      h_forcing + (size_t)w * SEMIC_NFORCING * ldh + p0, (size_t)ldh * 8, (size_t)np * 8,
                                      (size_t)run * SEMIC_NFORCING, cudaMemcpyHostToDevice, s_in));
                if (use_vali)
                    CUX(cudaMemcpy2DAsync(d_vs[b] + (size_t)done * SEMIC_NOUT * np, (size_t)np * 8,
                                          h_vali + (size_t)w * SEMIC_NOUT * ldh + p0, (size_t)ldh * 8, (size_t)np * 8,
                                          (size_t)run * SEMIC_NOUT, cudaMemcpyHostToDevice, s_in));
                done += run;
            }
            CUX(cudaEventRecord(ev_in[b], s_in));
            // compute stream: re-tile, run the unit, de-tile its output days
            CUX(cudaStreamWaitEvent(now->stream, ev_in[b], 0));
            k_soa_to_tiled<<<1184, 256, 0, now->stream>>>(d_fs[b], np, d_ft[b], sld, SEMIC_NFORCING, 0, nd, np);
            if (use_vali) k_soa_to_tiled<<<1184, 256, 0, now->stream>>>(d_vs[b], np, d_vt[b], sld, SEMIC_NOUT, 0, nd, np);
            CUX(cudaGetLastError());
            CUX(cudaEventRecord(ev_t[b], now->stream));
            if (o_n > 0 && unit >= 2) CUX(cudaStreamWaitEvent(now->stream, ev_o[b], 0));
            RunArgs a;
            fill_run_args(a, now, par, bnd, d_ft[b], nd, use_vali ? d_vt[b] : nullptr, 0, nd, d_ot[b], chunk_days, SEMIC_NOUT, o_n);
            a.slab = now->d_slab + p0; a.mask = now->d_mask + p0; a.nx = np; a.sld = sld;      // this slice of the points
            a.avg_days = avg;
            const int o_rows = (o_n + avg - 1) / avg;                         // rows this unit writes (= o_n for daily output)
            cudaError_t e = strict ? launch_run_strict(a, generic, now->stream, &launches_total)
                                   : launch_run_fast(a, generic, now->stream, &launches_total);
            if (e != cudaSuccess) { status = cuda_fail(e, "semic_run_host: kernel launch"); cleanup(); return status; }
            if (o_n > 0) {
                k_tiled_to_soa<<<1184, 256, 0, now->stream>>>(d_ot[b], sld, SEMIC_NOUT, 0, o_rows, 0, np, d_os[b], np);
                CUX(cudaGetLastError());
            }
            if (c == nchunks - 1) {           // the state's forcing fields hold the last day's forcing, as after the reference's loop
                k_tiled_day_to_slab<<<(np + 255) / 256, 256, 0, now->stream>>>(d_ft[b], sld, nd - 1, np, now->d_slab + p0, ld);
                CUX(cudaGetLastError());
            }
            CUX(cudaEventRecord(ev_k[b], now->stream));
            if (o_n > 0) {                                                    // D2H of the unit's output days
                CUX(cudaStreamWaitEvent(s_out, ev_k[b], 0));
                CUX(cudaMemcpy2DAsync(h_out + (size_t)((o_lo - out_first) / avg) * SEMIC_NOUT * ldh + p0, (size_t)ldh * 8, d_os[b],
                                      (size_t)np * 8, (size_t)np * 8, (size_t)o_rows * SEMIC_NOUT, cudaMemcpyDeviceToHost, s_out));
                CUX(cudaEventRecord(ev_o[b], s_out));
            }
        }
    }
    for (int b = 0; b < 2; ++b) if (out_days > 0 && unit > b) CUX(cudaStreamWaitEvent(now->stream, ev_o[b], 0));
    CUX(cudaEventRecord(t1, now->stream));
    CUX(cudaStreamSynchronize(now->stream));
    CUX(cudaStreamSynchronize(s_out));
    if (total_ms) CUX(cudaEventElapsedTime(total_ms, t0, t1));
#undef CUX
    cleanup();
    return SEMIC_OK;
}

// ------------------------------------------------------------------------------------------------
// averaging (surface_physics.f90:565-629)
// ------------------------------------------------------------------------------------------------
extern "C" int semic_surface_physics_average(semic_state *ave, const semic_state *now, const char *step, double nt)
{
    if (!ave || !now || !step || ave->npts != now->npts) { set_error("surface_physics_average: bad arguments"); return SEMIC_ERR_ARG; }
    const std::string st = ftrim(step, (int)std::strlen(step));
    int mode;
    if (st == "init") mode = 0;
    else if (st == "step") mode = 1;
    else if (st == "end") {
        mode = 2;
        if (!(nt >= 0.0)) { set_error("Averaging step total not provided."); return SEMIC_ERR_STEP; }      // :616-619
    } else { set_error("Step not recognized: %s", st.c_str()); return SEMIC_ERR_STEP; }                     // :623-624
    CU(cudaSetDevice(ave->device));
    if (ave->npts == 0) return SEMIC_OK;
    // order after work queued on the source state's stream
    CU(cudaEventRecord(now->ev1, now->stream));
    CU(cudaStreamWaitEvent(ave->stream, now->ev1, 0));
    k_average<<<(ave->npts + 255) / 256, 256, 0, ave->stream>>>(ave->d_slab, now->d_slab, ave->npts, ave->ld, now->ld, mode, nt);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ave->stream));
    return SEMIC_OK;
}

// ------------------------------------------------------------------------------------------------
// elementals on arrays (host pointers in, host pointers out)
// ------------------------------------------------------------------------------------------------
static int run_elemental(int which, int nin, const double *const *in, const int *mask, const double *sc, int nsc, int nout,
                         double *const *out, int n)
{
    if (n < 0) { set_error("elemental: n < 0"); return SEMIC_ERR_ARG; }
    if (n == 0) return SEMIC_OK;
    std::vector<double *> d_in(nin, nullptr), d_out(nout, nullptr);
    int *d_mask = nullptr;
    int status = SEMIC_OK;
    (void)nsc;
    auto cleanup = [&]() {
        for (auto p : d_in) if (p) cudaFree(p);
        for (auto p : d_out) if (p) cudaFree(p);
        if (d_mask) cudaFree(d_mask);
    };
#define CUX(call)                                                                  \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) { status = cuda_fail(e__, #call); cleanup(); return status; } \
    } while (0)
    for (int k = 0; k < nin; ++k) {
        if (!in[k]) { set_error("elemental: null input"); cleanup(); return SEMIC_ERR_ARG; }
        CUX(cudaMalloc(&d_in[k], (size_t)n * 8));
        CUX(cudaMemcpy(d_in[k], in[k], (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    for (int k = 0; k < nout; ++k) {
        if (!out[k]) { set_error("elemental: null output"); cleanup(); return SEMIC_ERR_ARG; }
        CUX(cudaMalloc(&d_out[k], (size_t)n * 8));
    }
    if (mask) { CUX(cudaMalloc(&d_mask, (size_t)n * 4)); CUX(cudaMemcpy(d_mask, mask, (size_t)n * 4, cudaMemcpyHostToDevice)); }
    CUX(launch_elementals_fast(which, d_in.data(), d_mask, sc, d_out.data(), n, nullptr));
    CUX(cudaDeviceSynchronize());
    for (int k = 0; k < nout; ++k) CUX(cudaMemcpy(out[k], d_out[k], (size_t)n * 8, cudaMemcpyDeviceToHost));
#undef CUX
    cleanup();
    return SEMIC_OK;
}

extern "C" int semic_sensible_heat_flux(const double *ts, const double *ta, const double *wind, const double *rhoa, double csh,
                                        double cap, double *shf, int n)
{
    const double *in[4] = {ts, ta, wind, rhoa};
    const double sc[2] = {csh, cap};
    double *out[1] = {shf};
    return run_elemental(0, 4, in, nullptr, sc, 2, 1, out, n);
}
extern "C" int semic_latent_heat_flux(const double *ts, const double *wind, const double *shum, const double *sp,
                                      const double *rhoatm, const int *mask, double clh, double eps, double cls, double clv,
                                      double *lhf, double *subl, double *evap, int n)
{
    const double *in[5] = {ts, wind, shum, sp, rhoatm};
    const double sc[4] = {clh, eps, cls, clv};
    double *out[3] = {lhf, subl, evap};
    return run_elemental(1, 5, in, mask, sc, 4, 3, out, n);
}
extern "C" int semic_longwave_upward(const double *ts, double sigma, double *lwout, int n)
{
    const double *in[1] = {ts};
    const double sc[1] = {sigma};
    double *out[1] = {lwout};
    return run_elemental(2, 1, in, nullptr, sc, 1, 1, out, n);
}

// one value as the function result (for Fortran ELEMENTAL specifics, see the header); NaN when the call fails
extern "C" double semic_sensible_heat_flux_1(double ts, double ta, double wind, double rhoa, double csh, double cap)
{
    double o = 0.0;
    return semic_sensible_heat_flux(&ts, &ta, &wind, &rhoa, csh, cap, &o, 1) == SEMIC_OK ? o : std::nan("");
}
extern "C" double semic_latent_heat_flux_1(double ts, double wind, double shum, double sp, double rhoatm, int mask, double clh,
                                           double eps, double cls, double clv, int which)
{
    double o[3] = {0.0, 0.0, 0.0};
    if (which < 0 || which > 2) { set_error("latent_heat_flux_1: which must be 0 (lhf), 1 (subl) or 2 (evap)"); return std::nan(""); }
    return semic_latent_heat_flux(&ts, &wind, &shum, &sp, &rhoatm, &mask, clh, eps, cls, clv, &o[0], &o[1], &o[2], 1) == SEMIC_OK
               ? o[which] : std::nan("");
}
extern "C" double semic_longwave_upward_1(double ts, double sigma)
{
    double o = 0.0;
    return semic_longwave_upward(&ts, sigma, &o, 1) == SEMIC_OK ? o : std::nan("");
}

// ------------------------------------------------------------------------------------------------
// NCCL (resolved lazily so that the library loads on machines without libnccl)
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load()
{
    if (g_nccl.h) return SEMIC_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) { set_error("NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror()); return SEMIC_ERR_NCCL; }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        set_error("NCCL: missing symbols in libnccl");
        return SEMIC_ERR_NCCL;
    }
    g_nccl.h = h;
    return SEMIC_OK;
}
int nccl_fail(ncclResult_t r, const char *what)
{
    set_error("%s: NCCL error %d (%s)", what, (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return SEMIC_ERR_NCCL;
}
}  // namespace

// Peer-memory set-up: allocate this rank's exchange buffer, hand its IPC handle to every rank (ncclAllGather of the 64 handle
// bytes), map the peers' buffers, and agree (ncclAllReduce, min) on whether EVERY rank succeeded -- the ranks must take the same
// path or they would wait for each other forever.  SEMIC_COMM_P2P=0 keeps NCCL's all-reduce (the baseline to compare with).
static void p2p_setup(semic_comm *c)
{
    auto off = [&](const char *why) { std::snprintf(c->p2p_why, sizeof c->p2p_why, "%s", why); };
    const char *env = getenv("SEMIC_COMM_P2P");
    int want = !(env && std::strcmp(env, "0") == 0) && g_nccl.AllGather != nullptr;
    const size_t bytes = 128 + 2 * (size_t)kP2PMaxPar * sizeof(double);
    int ok = want;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof mine);
    if (ok && (cudaMalloc(&c->p2p_own, bytes) != cudaSuccess || cudaMemset(c->p2p_own, 0, bytes) != cudaSuccess ||
               cudaDeviceSynchronize() != cudaSuccess || cudaIpcGetMemHandle(&mine, c->p2p_own) != cudaSuccess)) {
        ok = 0;
        cudaGetLastError();
    }
    // exchange handles and "ok so far" in one all-gather: [64 handle bytes][1 ok byte][pad] per rank
    const size_t rec = 72;
    std::vector<unsigned char> all(rec * (size_t)c->nranks, 0), me(rec, 0);
    std::memcpy(me.data(), &mine, sizeof mine);
    me[64] = (unsigned char)ok;
    unsigned char *d_me = nullptr, *d_all = nullptr;
    bool gathered = false;
    if (cudaMalloc(&d_me, rec) == cudaSuccess && cudaMalloc(&d_all, rec * (size_t)c->nranks) == cudaSuccess &&
        cudaMemcpy(d_me, me.data(), rec, cudaMemcpyHostToDevice) == cudaSuccess && g_nccl.AllGather &&
        g_nccl.AllGather(d_me, d_all, rec, ncclChar, c->comm, c->stream) == ncclSuccess &&
        cudaStreamSynchronize(c->stream) == cudaSuccess &&
        cudaMemcpy(all.data(), d_all, rec * (size_t)c->nranks, cudaMemcpyDeviceToHost) == cudaSuccess)
        gathered = true;
    if (d_me) cudaFree(d_me);
    if (d_all) cudaFree(d_all);
    if (!gathered) { off("handle exchange failed"); ok = 0; }
    for (int r = 0; ok && r < c->nranks; ++r) if (!all[rec * (size_t)r + 64]) ok = 0;      // somebody opted out or failed
    c->p2p_peer.assign((size_t)c->nranks, nullptr);
    for (int r = 0; ok && r < c->nranks; ++r) {
        if (r == c->rank) { c->p2p_peer[(size_t)r] = c->p2p_own; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, all.data() + rec * (size_t)r, sizeof h);
        if (cudaIpcOpenMemHandle(&c->p2p_peer[(size_t)r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            ok = 0;
            cudaGetLastError();
            off("cudaIpcOpenMemHandle failed (no peer access between the GPUs?)");
        }
    }
    if (ok) {
        std::vector<double *> hd((size_t)c->nranks);
        std::vector<unsigned long long *> hf((size_t)c->nranks);
        for (int r = 0; r < c->nranks; ++r) {
            hf[(size_t)r] = reinterpret_cast<unsigned long long *>(c->p2p_peer[(size_t)r]);
            hd[(size_t)r] = reinterpret_cast<double *>(reinterpret_cast<char *>(c->p2p_peer[(size_t)r]) + 128);
        }
        if (cudaMalloc(&c->d_data, sizeof(double *) * (size_t)c->nranks) != cudaSuccess ||
            cudaMalloc(&c->d_flags, sizeof(unsigned long long *) * (size_t)c->nranks) != cudaSuccess ||
            cudaMalloc(&c->d_done, sizeof(unsigned)) != cudaSuccess || cudaMalloc(&c->d_err, sizeof(int)) != cudaSuccess ||
            cudaMemcpy(c->d_data, hd.data(), sizeof(double *) * (size_t)c->nranks, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(c->d_flags, hf.data(), sizeof(unsigned long long *) * (size_t)c->nranks, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemset(c->d_done, 0, sizeof(unsigned)) != cudaSuccess || cudaMemset(c->d_err, 0, sizeof(int)) != cudaSuccess ||
            cudaDeviceSynchronize() != cudaSuccess) {
            ok = 0;
            cudaGetLastError();
            off("device tables");
        }
    }
    // agree: every rank must have mapped every peer
    {
        int *d_ok = nullptr;
        int h_ok = ok;
        bool agreed = false;
        if (cudaMalloc(&d_ok, sizeof(int)) == cudaSuccess && cudaMemcpy(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
            g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, c->stream) == ncclSuccess &&
            cudaStreamSynchronize(c->stream) == cudaSuccess && cudaMemcpy(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess)
            agreed = true;
        if (d_ok) cudaFree(d_ok);
        if (!agreed) h_ok = 0;
        if (ok && !h_ok) off("another rank could not map its peers");
        ok = h_ok;
    }
    if (!want && !c->p2p_why[0]) off(env ? "SEMIC_COMM_P2P=0" : "ncclAllGather not available");
    if (!ok && !c->p2p_why[0]) off("peer-memory set-up failed on a rank");
    c->p2p_ready = ok != 0;
}

// for semic_ensemble.cu: the arguments of the fused reduction + exchange kernel for the NEXT exchange, or false
namespace semic {
bool comm_p2p_next(semic_comm *c, int npar, P2PArgs *a)
{
    if (!c || !c->p2p_ready || npar > kP2PMaxPar) return false;
    ++c->epoch;
    a->data = c->d_data; a->flags = c->d_flags; a->done = c->d_done; a->err = c->d_err;
    a->nranks = c->nranks; a->rank = c->rank; a->epoch = c->epoch;
    a->timeout_cycles = 20LL * 1000 * 1000 * 1000;           // ~10 s at 2 GHz: a missing peer is an error, not a hang
    return true;
}
int comm_p2p_check(semic_comm *c, float ms)
{
    int err = 0;
    if (cudaMemcpy(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) err = 1;
    c->last_ms = ms;
    if (err) { set_error("peer-memory exchange: a rank did not arrive within the time-out"); return SEMIC_ERR_NCCL; }
    return SEMIC_OK;
}
}  // namespace semic

extern "C" int semic_comm_unique_id(char id[128])
{
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    ncclResult_t r = g_nccl.GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    std::memcpy(id, &u, 128);
    return SEMIC_OK;
}
extern "C" int semic_comm_init(semic_comm **out, const char id[128], int nranks, int rank)
{
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) { set_error("comm_init: bad arguments"); return SEMIC_ERR_ARG; }
    int rc = nccl_load();
    if (rc) return rc;
    semic_comm *c = new semic_comm();
    c->nranks = nranks; c->rank = rank;
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_nccl.CommDestroy(c->comm); delete c; return cuda_fail(e, "comm_init: stream"); }
    if ((e = cudaEventCreate(&c->e0)) != cudaSuccess || (e = cudaEventCreate(&c->e1)) != cudaSuccess) {
        g_nccl.CommDestroy(c->comm); cudaStreamDestroy(c->stream); delete c; return cuda_fail(e, "comm_init: events");
    }
    cudaGetDevice(&c->device);
    p2p_setup(c);                                 // optional: without it the exchange is NCCL's all-reduce
    *out = c;
    return SEMIC_OK;
}
extern "C" int semic_comm_p2p(const semic_comm *c, char *why, int why_len)
{
    if (!c) return 0;
    if (why && why_len > 0) std::snprintf(why, (size_t)why_len, "%s", c->p2p_why);
    return c->p2p_ready ? 1 : 0;
}
extern "C" int semic_comm_allreduce_sum(semic_comm *c, double *dev_buf, int count)
{
    if (!c || !dev_buf || count < 0) { set_error("comm_allreduce_sum: bad arguments"); return SEMIC_ERR_ARG; }
    CU(cudaEventRecord(c->e0, c->stream));
    ncclResult_t r = g_nccl.AllReduce(dev_buf, dev_buf, (size_t)count, ncclFloat64, ncclSum, c->comm, c->stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
    CU(cudaEventRecord(c->e1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&c->last_ms, c->e0, c->e1));
    return SEMIC_OK;
}
extern "C" int semic_comm_last_allreduce_ms(const semic_comm *c, float *ms)
{
    if (!c || !ms) { set_error("comm_last_allreduce_ms: bad arguments"); return SEMIC_ERR_ARG; }
    *ms = c->last_ms;
    return SEMIC_OK;
}
extern "C" int semic_comm_free(semic_comm *c)
{
    if (!c) { set_error("comm_free: null"); return SEMIC_ERR_ARG; }
    cudaSetDevice(c->device);
    if (c->p2p_ready) {
        // collective: nobody unmaps or frees while a peer may still be reading (one tiny all-reduce as the barrier)
        int *d_one = nullptr;
        if (cudaMalloc(&d_one, sizeof(int)) == cudaSuccess && cudaMemset(d_one, 0, sizeof(int)) == cudaSuccess) {
            g_nccl.AllReduce(d_one, d_one, 1, ncclInt32, ncclSum, c->comm, c->stream);
            cudaStreamSynchronize(c->stream);
        }
        if (d_one) cudaFree(d_one);
    }
    for (int r = 0; r < (int)c->p2p_peer.size(); ++r)
        if (c->p2p_peer[(size_t)r] && c->p2p_peer[(size_t)r] != c->p2p_own) cudaIpcCloseMemHandle(c->p2p_peer[(size_t)r]);
    if (c->d_data) cudaFree(c->d_data);
    if (c->d_flags) cudaFree(c->d_flags);
    if (c->d_done) cudaFree(c->d_done);
    if (c->d_err) cudaFree(c->d_err);
    if (c->p2p_own) cudaFree(c->p2p_own);
    if (c->e0) cudaEventDestroy(c->e0);
    if (c->e1) cudaEventDestroy(c->e1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return SEMIC_OK;
}

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
extern "C" int semic_set_device(int device) { CU(cudaSetDevice(device)); return SEMIC_OK; }
extern "C" int semic_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
extern "C" int semic_device_synchronize(void) { CU(cudaDeviceSynchronize()); return SEMIC_OK; }
extern "C" const char *semic_last_error(void) { return g_err; }
extern "C" const char *semic_version(void) { return "semic_b200 0.1 (sm_100a)"; }
extern "C" int semic_kernel_launches(void) { return launches_total; }

extern "C" int semic_probe_dfma(int iters, float *ms, double *flops)
{
    int dev = 0, nsm = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = nsm * 8, threads = 256;
    double *out = nullptr;
    CU(cudaMalloc(&out, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    k_probe_dfma<<<blocks, threads>>>(out, 16);
    CU(cudaEventRecord(e0));
    k_probe_dfma<<<blocks, threads>>>(out, iters);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    CU(cudaGetLastError());
    if (ms) CU(cudaEventElapsedTime(ms, e0, e1));
    if (flops) *flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return SEMIC_OK;
}

extern "C" int semic_probe_copy(size_t bytes, float *ms)
{
    int dev = 0, nsm = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    bytes &= ~(size_t)15;
    double2 *a = nullptr, *b = nullptr;
    CU(cudaMalloc(&a, bytes));
    CU(cudaMalloc(&b, bytes));
    CU(cudaMemset(a, 1, bytes));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    k_probe_copy<<<nsm * 16, 512>>>(a, b, bytes / 16);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CU(cudaEventRecord(e0));
        k_probe_copy<<<nsm * 16, 512>>>(a, b, bytes / 16);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float t;
        CU(cudaEventElapsedTime(&t, e0, e1));
        if (t < best) best = t;
    }
    CU(cudaGetLastError());
    if (ms) *ms = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    return SEMIC_OK;
}

extern "C" int semic_flush_l2(void)
{
    static double *buf = nullptr;
    static int buf_dev = -1;
    const size_t bytes = 256ull << 20;                                   // 256 MiB > 126 MB L2
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (!buf || buf_dev != dev) {
        CU(cudaMalloc(&buf, bytes));
        buf_dev = dev;
    }
    CU(cudaMemset(buf, 0, bytes));
    CU(cudaDeviceSynchronize());
    return SEMIC_OK;
}
