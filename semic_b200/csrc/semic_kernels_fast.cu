// Fast arithmetic instantiation (the benchmarked path): FMA contraction, reciprocal multiplies,
// MUFU-seeded division, custom exp.  Parity bar: 1e-10 relative per variable after one year.
#define SEMIC_STRICT 0
#define SEMIC_MODE_NS fast_mode
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "semic_kernels.cuh"

namespace semic {
namespace fast_mode {
#include "semic_launch.inl"

// public elementals evaluated over arrays (surface_physics.f90:378-438)
__global__ void k_elementals(int which, const double *in0, const double *in1, const double *in2, const double *in3,
                             const double *in4, const int *mask, double s0, double s1, double s2, double s3,
                             double *o0, double *o1, double *o2, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    (void)mask;                                            // unused by the reference as well (:404)
    if (which == 0) {
        o0[i] = sensible_heat_flux(in0[i], in1[i], in2[i], in3[i], s0, s1);
    } else if (which == 1) {
        const double ts = in0[i], wind = in1[i], shum = in2[i], sp = in3[i], rho = in4[i];
        const bool ice = ts < c_t0;
        const double es = esat(ts, ice);
        const double shum_sat = fdiv(es * s1, (es * (s1 - 1.0)) + sp);
        const double q = ((s0 * wind) * rho) * (shum_sat - shum);
        o0[i] = q * (ice ? s2 : s3);
        o1[i] = ice ? q : 0.0;
        o2[i] = ice ? 0.0 : q;
    } else {
        o0[i] = longwave_upward(in0[i], s0);
    }
}
}  // namespace fast_mode

// g_exp2_tab of the current device: 2^(j/2048) from the host's long-double exp2 (x87: 64-bit mantissa), rounded to double
static cudaError_t ensure_exp_table()
{
    constexpr int kMaxDev = 64;
    static std::mutex mu;
    static bool done[kMaxDev] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDev) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return cudaSuccess;
    static double host_tab[fast_mode::kExpTab];
    for (int j = 0; j < fast_mode::kExpTab; ++j) host_tab[j] = (double)exp2l((long double)j / (long double)fast_mode::kExpTab);
    // a pageable-memory copy on the legacy stream may return before the DMA has landed, and the kernels that read the
    // table run on non-blocking streams (not ordered behind the legacy stream): wait for it before declaring it done
    e = cudaMemcpyToSymbol(fast_mode::g_exp2_tab, host_tab, sizeof host_tab);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) done[dev] = true;
    return e;
}

cudaError_t prepare_fast() { return ensure_exp_table(); }

cudaError_t launch_run_fast(const RunArgs &a, bool generic, cudaStream_t st, int *launches)
{
    if (launches) ++*launches;
    if (cudaError_t e = ensure_exp_table()) return e;
    const char *e = getenv("SEMIC_SHAPE");
    const bool force_general = e && std::strcmp(e, "24x1") == 0;
    const bool lean_ok = a.f_tiled && a.ndays >= 1 && a.P.n_ksub >= 1 && (a.out == nullptr || a.o_nvar == SEMIC_NOUT);
    const bool want_avg = a.out != nullptr && a.avg_days > 1;
    if (generic) {
        // BASELINE configs[3]: only tsurf and/or alb are overridden, from an external series -> the plain path with two
        // more rows per stage; every other flag set runs the general kernel (all 23 fields carried)
        const semic_bnd &B = a.B;
        const bool only_ts_alb = (B.tsurf || B.alb) && !(B.shf || B.lhf || B.amp || B.melt || B.refr || B.acc || B.smb);
        if (only_ts_alb && a.vali != nullptr && lean_ok && !want_avg && !force_general) return fast_mode::launch_lean(a, st, 1);
        // the other flag sets (melt / acc / smb from the external series; shf lhf amp refr frozen at their state values): the
        // GENERIC physics on the same shell.  k_run<GENERIC> keeps the per-day call, energy_balance / mass_balance alone, the bitmap
        if (lean_ok && !want_avg && !force_general && a.do_mb && a.eb_steps == a.P.n_ksub) return fast_mode::launch_lean(a, st, 2);
        return fast_mode::launch_scheme<true, 1, 8>(a, st);
    }
    // Plain runs go to k_run_lean; what it does not cover (the bitmap row, the drop-in single day on the slab's own
    // forcing fields, n_ksub < 1) runs k_run.  SEMIC_SHAPE=24x1 forces k_run for A/B timing (tools/ab.py).
    if (want_avg && !lean_ok) return cudaErrorNotSupported;                  // period means: plain runs only
    if (lean_ok && (want_avg || !force_general)) return fast_mode::launch_lean(a, st, 0);
    return a.P.n_ksub >= 8 ? fast_mode::launch_scheme<false, 1, 24, 4>(a, st) : fast_mode::launch_scheme<false, 1, 24, 2>(a, st);
}

cudaError_t launch_vali_invvar(const EnsArgs &a, double *inv_var, cudaStream_t st)
{
    if (a.nx <= 0) return cudaSuccess;
    fast_mode::k_vali_invvar<<<(a.nx + 127) / 128, 128, 0, st>>>(a.vali, a.ntime, a.nx, a.ld, inv_var);
    return cudaGetLastError();
}

cudaError_t launch_ensemble_fast(const EnsArgs &a, double *sumsq_dev, cudaStream_t st, const P2PArgs *p2p, cudaEvent_t mid)
{
    if (a.npar <= 0) return cudaSuccess;
    if (a.nx <= 0) {                                    // a rank without points still takes part in the exchange (zeros)
        if (!p2p) return cudaSuccess;
        if (mid) { if (cudaError_t e = cudaEventRecord(mid, st)) return e; }
        fast_mode::k_ens_reduce_p2p<<<(a.npar + 127) / 128, 128, 0, st>>>(a.partial, 0, a.npar, sumsq_dev, *p2p);
        return cudaGetLastError();
    }
    if (cudaError_t e = ensure_exp_table()) return e;
    const int ngroups = (a.npar + fast_mode::kEnsParticles - 1) / fast_mode::kEnsParticles;
    const long long nblocks = (long long)a.ntiles * ngroups;
    if (nblocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)nblocks), block(32 * fast_mode::kEnsParticles);
    const size_t smem = fast_mode::EnsSmem::bytes;
    const bool k3 = a.uni_ksub == 3;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kern<<<grid, block, smem, st>>>(a, ngroups);
        return cudaGetLastError();
    };
    cudaError_t e;
    if (a.uni_scheme == SCHEME_NONE) e = k3 ? go(fast_mode::k_ensemble<SCHEME_NONE, 3>) : go(fast_mode::k_ensemble<SCHEME_NONE, 0>);
    else                             e = k3 ? go(fast_mode::k_ensemble<SCHEME_RUNTIME, 3>) : go(fast_mode::k_ensemble<SCHEME_RUNTIME, 0>);
    if (e != cudaSuccess) return e;
    if (mid) { if ((e = cudaEventRecord(mid, st)) != cudaSuccess) return e; }
    if (p2p) fast_mode::k_ens_reduce_p2p<<<(a.npar + 127) / 128, 128, 0, st>>>(a.partial, a.ntiles, a.npar, sumsq_dev, *p2p);
    else     fast_mode::k_ens_reduce<<<(a.npar + 127) / 128, 128, 0, st>>>(a.partial, a.ntiles, a.npar, sumsq_dev);
    return cudaGetLastError();
}

cudaError_t launch_elementals_fast(int which, const double *const *in, const int *mask, const double *sc,
                                   double *const *out, int n, cudaStream_t st)
{
    const double *i0 = in[0], *i1 = nullptr, *i2 = nullptr, *i3 = nullptr, *i4 = nullptr;
    double *o0 = out[0], *o1 = nullptr, *o2 = nullptr;
    double s0 = sc[0], s1 = 0, s2 = 0, s3 = 0;
    if (which == 0) { i1 = in[1]; i2 = in[2]; i3 = in[3]; s1 = sc[1]; }
    if (which == 1) { i1 = in[1]; i2 = in[2]; i3 = in[3]; i4 = in[4]; s1 = sc[1]; s2 = sc[2]; s3 = sc[3]; o1 = out[1]; o2 = out[2]; }
    fast_mode::k_elementals<<<(n + 255) / 256, 256, 0, st>>>(which, i0, i1, i2, i3, i4, mask, s0, s1, s2, s3, o0, o1, o2, n);
    return cudaGetLastError();
}
}  // namespace semic
