// semic_ensemble.cu -- C-ABI entry of the parameter-ensemble misfit (optimize/run_particles.f90:104-178).
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "semic_b200.h"
#include "semic_internal.h"

namespace semic {
void set_error(const char *fmt, ...);
DevPar make_devpar_public(const semic_par *par);
bool any_consulted_flag_public(const semic_bnd *b);
int state_info(const semic_state *s, int *npts, int *ld, int *device, double **d_slab, int **d_mask, cudaStream_t *stream);
int series_info(const semic_series *s, int *npts, int *ld, int *nvar, int *ndays, double **d);
bool comm_p2p_next(semic_comm *c, int npar, P2PArgs *a);
int comm_p2p_check(semic_comm *c, float ms);
}  // namespace semic
using namespace semic;

#define CUE(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            set_error("%s: CUDA error %d (%s)", #call, (int)e__, cudaGetErrorString(e__));              \
            status = SEMIC_ERR_CUDA;                                                                    \
            goto done;                                                                                  \
        }                                                                                               \
    } while (0)

extern "C" int semic_ensemble_cost(const semic_state *init, const semic_par *pars, int npar, const semic_bnd *bnd,
                                   const semic_series *forcing, const semic_series *vali, int ntime, int nloop,
                                   semic_comm *comm, double *sumsq, float *kernel_ms)
{
    int status = SEMIC_OK;
    int npts = 0, ld = 0, device = 0, f_npts = 0, f_ld = 0, f_nvar = 0, f_ndays = 0, v_npts = 0, v_ld = 0, v_nvar = 0, v_ndays = 0;
    double *d_slab = nullptr, *d_f = nullptr, *d_v = nullptr;
    int *d_mask = nullptr;
    cudaStream_t st = nullptr;
    double *d_invvar = nullptr, *d_partial = nullptr, *d_sumsq = nullptr;
    DevPar *d_pars = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, emid = nullptr;
    std::vector<DevPar> hp;
    EnsArgs a;
    P2PArgs p2p;
    bool use_p2p = false;

    if (!init || !pars || npar <= 0 || !bnd || !forcing || !vali || ntime <= 0 || nloop <= 0 || !sumsq) {
        set_error("semic_ensemble_cost: bad arguments");
        return SEMIC_ERR_ARG;
    }
    if (any_consulted_flag_public(bnd)) {
        set_error("semic_ensemble_cost: boundary overrides are not supported in the ensemble path "
                  "(optimize/tools.py:9 always writes boundary = \"\", \"\", \"\")");
        return SEMIC_ERR_ARG;
    }
    state_info(init, &npts, &ld, &device, &d_slab, &d_mask, &st);
    series_info(forcing, &f_npts, &f_ld, &f_nvar, &f_ndays, &d_f);
    series_info(vali, &v_npts, &v_ld, &v_nvar, &v_ndays, &d_v);
    if (f_nvar != SEMIC_NFORCING || v_nvar != SEMIC_NOUT || f_npts != npts || v_npts != npts || f_ndays < ntime ||
        v_ndays < ntime) {
        set_error("semic_ensemble_cost: forcing must be 9 x npts x >=ntime, validation 8 x npts x >=ntime");
        return SEMIC_ERR_ARG;
    }
    for (int p = 0; p < npar; ++p) {
        if (pars[p].n_ksub < 1) { set_error("semic_ensemble_cost: n_ksub < 1 in parameter set %d", p); return SEMIC_ERR_ARG; }
        hp.push_back(make_devpar_public(&pars[p]));
    }
    if (kernel_ms) *kernel_ms = 0.f;
    for (int p = 0; p < npar; ++p) sumsq[p] = 0.0;

    std::memset(&a, 0, sizeof a);
    CUE(cudaSetDevice(device));
    a.ntiles = (npts + 31) / 32;
    if (npts > 0) {
        CUE(cudaMalloc(&d_invvar, (size_t)4 * ld * sizeof(double)));
        CUE(cudaMemsetAsync(d_invvar, 0, (size_t)4 * ld * sizeof(double), st));   // on the stream the kernels run on
        CUE(cudaMalloc(&d_partial, (size_t)npar * a.ntiles * sizeof(double)));
    }
    CUE(cudaMalloc(&d_sumsq, (size_t)npar * sizeof(double)));
    CUE(cudaMemsetAsync(d_sumsq, 0, (size_t)npar * sizeof(double), st));
    CUE(cudaMalloc(&d_pars, (size_t)npar * sizeof(DevPar)));
    CUE(cudaMemcpyAsync(d_pars, hp.data(), (size_t)npar * sizeof(DevPar), cudaMemcpyHostToDevice, st));
    CUE(cudaStreamSynchronize(st));                // hp is pageable: staged and landed before the launch below
    CUE(cudaEventCreate(&e0));
    CUE(cudaEventCreate(&e1));
    a.slab0 = d_slab; a.mask = d_mask; a.nx = npts; a.ld = ld;
    a.forcing = d_f; a.vali = d_v;
    if (f_ld != ld || v_ld != ld) { set_error("semic_ensemble_cost: leading dimensions differ"); status = SEMIC_ERR_ARG; goto done; }
    a.inv_var = d_invvar; a.pars = d_pars; a.npar = npar; a.ntime = ntime; a.nloop = nloop;
    a.partial = d_partial;
    a.uni_scheme = hp[0].scheme; a.uni_ksub = hp[0].n_ksub;
    for (int p = 1; p < npar; ++p) {
        if (hp[p].scheme != a.uni_scheme) a.uni_scheme = -1;
        if (hp[p].n_ksub != a.uni_ksub) a.uni_ksub = 0;
    }
    // The one collective of the path (sum over the shards).  With a peer-memory communicator it is part of the second
    // reduction stage (k_ens_reduce_p2p: every rank adds the ranks' sums over NVLink in rank order); otherwise NCCL's all-reduce.
    use_p2p = comm != nullptr && comm_p2p_next(comm, npar, &p2p);
    CUE(cudaEventCreate(&emid));
    if (npts > 0 || use_p2p) {
        if (npts > 0) CUE(launch_vali_invvar(a, d_invvar, st));
        CUE(cudaEventRecord(e0, st));
        CUE(launch_ensemble_fast(a, d_sumsq, st, use_p2p ? &p2p : nullptr, emid));
        CUE(cudaEventRecord(e1, st));
        CUE(cudaStreamSynchronize(st));
        if (use_p2p) {                            // kernel_ms = the trajectories; the reduction + exchange is the communicator's figure
            float ex = 0.f;
            if (kernel_ms) CUE(cudaEventElapsedTime(kernel_ms, e0, emid));
            CUE(cudaEventElapsedTime(&ex, emid, e1));
            const int rc = comm_p2p_check(comm, ex);
            if (rc) { status = rc; goto done; }
        } else if (kernel_ms) {
            CUE(cudaEventElapsedTime(kernel_ms, e0, e1));
        }
    }
    if (comm && !use_p2p) {
        const int rc = semic_comm_allreduce_sum(comm, d_sumsq, npar);
        if (rc) { status = rc; goto done; }
    }
    CUE(cudaMemcpy(sumsq, d_sumsq, (size_t)npar * sizeof(double), cudaMemcpyDeviceToHost));
done:
    if (d_invvar) cudaFree(d_invvar);
    if (d_partial) cudaFree(d_partial);
    if (d_sumsq) cudaFree(d_sumsq);
    if (d_pars) cudaFree(d_pars);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (emid) cudaEventDestroy(emid);
    return status;
}
