// probe_issue.cu -- how many dispatch cycles does an FP64 warp instruction cost an SM sub-partition, and are the
// integer/logic instructions next to it free?  Each kernel runs 32 warps per SM (8 per scheduler) of fully
// independent instruction streams: 8 DFMA chains per thread, optionally interleaved with independent 32-bit IMADs.
// Reported: SM cycles per DFMA warp-instruction per scheduler (elapsed cycles * 4 schedulers / DFMAs per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_issue tools/probe_issue.cu && ./probe_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, bool ALU>
__global__ void __launch_bounds__(1024, 1) k_probe(double *out, int *iout, int iters, long long *cycles)
{
    double a[8];
    int b[8];
    const double m = 1.0000001, c = 1e-9;
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = threadIdx.x * 1e-3 + j; b[j] = threadIdx.x + j; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = fma(a[j], m, c);
#pragma unroll
            for (int q = 0; q < NI; ++q) {
                if (ALU) asm volatile("xor.b32 %0, %0, 0x55;" : "+r"(b[(j + q) & 7]));          // LOP3, ALU pipe
                else     asm volatile("mad.lo.s32 %0, %0, 3, 1;" : "+r"(b[(j + q) & 7]));       // IMAD, FMA pipe
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
    int si = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += a[j]; si += b[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = si;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NI, bool ALU>
static void run(const char *label)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out; int *iout; long long *cyc;
    cudaMalloc(&out, (size_t)nsm * 1024 * 8); cudaMalloc(&iout, (size_t)nsm * 1024 * 4); cudaMalloc(&cyc, nsm * 8);
    const int iters = 20000;
    k_probe<NI, ALU><<<nsm, 1024>>>(out, iout, 100, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_probe<NI, ALU><<<nsm, 1024>>>(out, iout, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
    const double dfma_per_sched = 8.0 * iters * 8.0;             // 8 warps per scheduler x 8 DFMA per iteration
    printf("%-34s %8.3f ms   %.3f cycles per DFMA per scheduler   (%.3f per instruction)\n", label, ms, mean / dfma_per_sched,
           mean / dfma_per_sched / (1 + NI));
    cudaFree(out); cudaFree(iout); cudaFree(cyc);
}

int main()
{
    run<0, false>("DFMA only");
    run<1, false>("DFMA + 1 independent IMAD");
    run<2, false>("DFMA + 2 independent IMAD");
    run<3, false>("DFMA + 3 independent IMAD");
    run<1, true>("DFMA + 1 independent LOP3");
    run<2, true>("DFMA + 2 independent LOP3");
    run<3, true>("DFMA + 3 independent LOP3");
    return 0;
}
