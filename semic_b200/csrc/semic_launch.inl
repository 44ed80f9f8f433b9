// semic_launch.inl -- host-side launcher of k_run for the arithmetic mode of the including TU.
// (included inside namespace semic::SEMIC_MODE_NS)

template <int SCHEME, bool GENERIC, int PPT, int MINB>
static cudaError_t launch_one(const RunArgs &a, cudaStream_t st)
{
    auto kern = k_run<SCHEME, GENERIC, PPT, MINB>;
    constexpr size_t smem = RunSmem<GENERIC, PPT>::bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 0, occ = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem)) != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    constexpr int T = RunSmem<GENERIC, PPT>::T;
    const int ntiles = (a.nx + T - 1) / T;
    int grid = nsm * occ;                       // persistent: one CTA per resident slot, each walks its tiles
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return cudaSuccess;
    kern<<<grid, kThreads, smem, st>>>(a);
    return cudaGetLastError();
}

template <bool GENERIC, int PPT, int MINB>
static cudaError_t launch_scheme(const RunArgs &a, cudaStream_t st)
{
    if (GENERIC) return launch_one<SCHEME_RUNTIME, GENERIC, PPT, MINB>(a, st);      // scheme resolved at run time
    switch (a.P.scheme) {
        case SCHEME_NONE:   return launch_one<SCHEME_NONE, GENERIC, PPT, MINB>(a, st);
        case SCHEME_SLATER: return launch_one<SCHEME_SLATER, GENERIC, PPT, MINB>(a, st);
        case SCHEME_DENBY:  return launch_one<SCHEME_DENBY, GENERIC, PPT, MINB>(a, st);
        case SCHEME_ISBA:   return launch_one<SCHEME_ISBA, GENERIC, PPT, MINB>(a, st);
        case SCHEME_ALEX:   return launch_one<SCHEME_ALEX, GENERIC, PPT, MINB>(a, st);
        default:            return launch_one<SCHEME_UNKNOWN, GENERIC, PPT, MINB>(a, st);
    }
}
