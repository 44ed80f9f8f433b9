// probe_mix.cu -- does the INSTRUCTION MIX of the sub-step loop alone (no dependencies between instructions of different
// chains) run at the FP64 pipe's rate?  One iteration issues, per thread, 32 DFMA on 8 independent chains plus the other
// instructions of one sub-step in k_run_lean's proportion (29 FP64 : 8 ALU : 4 IMAD : 2 MUFU.RCP64H : 1 LDS), all independent.
// Variants add the groups one at a time.  Reported: SM cycles per iteration per warp and scheduler; the FP64 pipe alone
// needs 64 (32 DFMA x 2 cycles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_mix tools/probe_mix.cu && ./probe_mix [warps per SM, default 24]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__host__ __device__ constexpr int every(int n) { return n > 0 ? 32 / n : 64; }      // one of the n extra instructions every so many DFMA

template <int NALU, int NIMAD, int NMUFU, int NLDS>
__global__ void __launch_bounds__(1024, 1) k_mix(double *out, int iters, long long *cycles)
{
    __shared__ double sh[1024];
    sh[threadIdx.x] = threadIdx.x * 1e-3;
    __syncthreads();
    double a[8], r[2] = {1.5 + threadIdx.x * 1e-6, 2.5 + threadIdx.x * 1e-6}, l[2] = {0.0, 0.0};
    int b[8], ib[4];
    const double m = 1.0000001, c = 1e-9;
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(&sh[threadIdx.x]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = threadIdx.x * 1e-3 + j; b[j] = threadIdx.x + j; }
#pragma unroll
    for (int j = 0; j < 4; ++j) ib[j] = threadIdx.x * 3 + j;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = fma(a[j], m, c);
                const int n = q * 8 + j;                         // position in the iteration: spread the others evenly
                if (NALU > 0 && n % every(NALU) == 0 && n / every(NALU) < NALU)
                    asm volatile("xor.b32 %0, %0, 0x55;" : "+r"(b[(n / every(NALU)) & 7]));
                if (NIMAD > 0 && n % every(NIMAD) == 1 % every(NIMAD) && n / every(NIMAD) < NIMAD)
                    asm volatile("mad.lo.s32 %0, %0, 3, 1;" : "+r"(ib[(n / every(NIMAD)) & 3]));
                if (NMUFU > 0 && n % every(NMUFU) == 2 % every(NMUFU) && n / every(NMUFU) < NMUFU)
                    asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r[(n / every(NMUFU)) & 1]));
                if (NLDS > 0 && n % every(NLDS) == 3 % every(NLDS) && n / every(NLDS) < NLDS)
                    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(l[(n / every(NLDS)) & 1]) : "r"(saddr) : "memory");
            }
        }
    }
    const long long t1 = clock64();
    double s = r[0] + r[1] + l[0] + l[1];
    int si = ib[0] + ib[1] + ib[2] + ib[3];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += a[j]; si += b[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + si;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int g_threads = 768;

template <int NALU, int NIMAD, int NMUFU, int NLDS>
static void run(const char *label)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out; long long *cyc;
    cudaMalloc(&out, (size_t)nsm * 1024 * 8); cudaMalloc(&cyc, nsm * 8);
    const int iters = 5000;
    k_mix<NALU, NIMAD, NMUFU, NLDS><<<nsm, g_threads>>>(out, 100, cyc);
    cudaDeviceSynchronize();
    k_mix<NALU, NIMAD, NMUFU, NLDS><<<nsm, g_threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
    const double warps_per_sched = g_threads / 32 / 4.0;
    printf("%-52s %7.2f cycles per iteration per warp and scheduler (FP64 pipe alone: 64)\n", label, mean / iters / warps_per_sched);
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char **argv)
{
    if (argc > 1) g_threads = atoi(argv[1]) * 32;
    printf("%d warps per SM\n", g_threads / 32);
    run<0, 0, 0, 0>("32 DFMA");
    run<8, 0, 0, 0>("32 DFMA + 8 ALU (LOP3)");
    run<0, 4, 0, 0>("32 DFMA + 4 IMAD");
    run<0, 0, 2, 0>("32 DFMA + 2 MUFU.RCP64H");
    run<0, 0, 0, 1>("32 DFMA + 1 LDS.64");
    run<8, 4, 0, 0>("32 DFMA + 8 ALU + 4 IMAD");
    run<8, 4, 2, 1>("32 DFMA + 8 ALU + 4 IMAD + 2 MUFU + 1 LDS (sub-step mix)");
    run<16, 4, 2, 1>("32 DFMA + 16 ALU + 4 IMAD + 2 MUFU + 1 LDS");
    return 0;
}
