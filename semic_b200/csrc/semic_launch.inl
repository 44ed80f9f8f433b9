// semic_launch.inl -- host-side launcher of k_run for the arithmetic mode of the including TU.
// (included inside namespace semic::SEMIC_MODE_NS)

template <int SCHEME, bool GENERIC, int MINB, bool BR, int WPC, int UNR = 2>
static cudaError_t launch_one(const RunArgs &a, cudaStream_t st)
{
    auto kern = k_run<SCHEME, GENERIC, MINB, BR, WPC, UNR>;
    constexpr size_t smem = RunSmem<GENERIC, WPC>::bytes;
    constexpr int kThreads = WPC * 32;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 0, occ = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem)) != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    const long long ntiles = ((long long)a.nx + kTile - 1) / kTile;
    long long grid = (long long)nsm * occ;      // persistent: one CTA per resident slot, every warp walks its tiles
    const long long need = (ntiles + WPC - 1) / WPC;
    if (grid > need) grid = need;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(a);
    return cudaGetLastError();
}

// The branch bitmap (9th output row, flip report) costs ~20 instructions per point-day, so it is a separate
// instantiation; it uses the run-time scheme switch to keep the number of kernels down.
template <bool GENERIC, int MINB, int WPC, int UNR = 2>
static cudaError_t launch_scheme(const RunArgs &a, cudaStream_t st)
{
    if (a.out != nullptr && a.o_nvar > SEMIC_NOUT) return launch_one<SCHEME_RUNTIME, GENERIC, 1, true, 8>(a, st);
    if (GENERIC) return launch_one<SCHEME_RUNTIME, GENERIC, MINB, false, WPC>(a, st);      // scheme resolved at run time
    switch (a.P.scheme) {
        case SCHEME_NONE:   return launch_one<SCHEME_NONE, GENERIC, MINB, false, WPC, UNR>(a, st);
        case SCHEME_SLATER: return launch_one<SCHEME_SLATER, GENERIC, MINB, false, WPC, UNR>(a, st);
        case SCHEME_DENBY:  return launch_one<SCHEME_DENBY, GENERIC, MINB, false, WPC, UNR>(a, st);
        case SCHEME_ISBA:   return launch_one<SCHEME_ISBA, GENERIC, MINB, false, WPC, UNR>(a, st);
        case SCHEME_ALEX:   return launch_one<SCHEME_ALEX, GENERIC, MINB, false, WPC, UNR>(a, st);
        default:            return launch_one<SCHEME_UNKNOWN, GENERIC, MINB, false, WPC, UNR>(a, st);
    }
}

#if !SEMIC_STRICT
// k_run_lean: the plain (no boundary flag) tiled-forcing run with 8 daily rows, n_ksub >= 1
template <int SCHEME, int WPC, int UNR, int NK, bool AVG, int EXT>
static cudaError_t launch_lean_one(const RunArgs &a, cudaStream_t st)
{
    auto kern = k_run_lean<SCHEME, WPC, UNR, NK, AVG, EXT>;
    constexpr size_t smem = LeanSmem<WPC, EXT>::bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const long long ntiles = ((long long)a.nx + kTile - 1) / kTile;
    // persistent: one CTA per SM, every warp walks its tiles.  A small grid is SPREAD over all SMs (tile t on CTA t % grid)
    // instead of packed into full CTAs when that leaves every scheduler at least one warp short of full: a tile's run is one
    // dependent chain, and the fewer warps share a scheduler the faster each chain goes (60 000 points x 365 days, n_ksub = 24:
    // 1.92 ms spread, 2.63 ms packed into 79 CTAs).  Between WPC - 4 and WPC tiles per SM the fullest scheduler is full either
    // way and packing is as fast (n_ksub = 24) or faster (n_ksub = 3: 0.84 against 0.90 ms at 100 000 points).
    // SEMIC_SPREAD=0 always packs, SEMIC_SPREAD=1 always spreads (A/B timing).
    long long grid = nsm;
    long long need = (ntiles + WPC - 1) / WPC;
    {
        const char *sp = getenv("SEMIC_SPREAD");
        const bool spread = sp ? atoi(sp) != 0 : ntiles <= (long long)(WPC - 4) * nsm;
        if (spread) need = ntiles;
    }
    if (grid > need) grid = need;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, WPC * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <int WPC, int UNR, int NK, bool AVG, int EXT = 0>
static cudaError_t launch_lean_nk(const RunArgs &a, cudaStream_t st)
{
    switch (a.P.scheme) {
        case SCHEME_NONE:   return launch_lean_one<SCHEME_NONE, WPC, UNR, NK, AVG, EXT>(a, st);
        case SCHEME_SLATER: return launch_lean_one<SCHEME_SLATER, WPC, UNR, NK, AVG, EXT>(a, st);
        case SCHEME_DENBY:  return launch_lean_one<SCHEME_DENBY, WPC, UNR, NK, AVG, EXT>(a, st);
        case SCHEME_ISBA:   return launch_lean_one<SCHEME_ISBA, WPC, UNR, NK, AVG, EXT>(a, st);
        case SCHEME_ALEX:   return launch_lean_one<SCHEME_ALEX, WPC, UNR, NK, AVG, EXT>(a, st);
        default:            return launch_lean_one<SCHEME_UNKNOWN, WPC, UNR, NK, AVG, EXT>(a, st);
    }
}

// n_ksub = 3 (every shipped namelist: example.namelist:30, optimization.namelist, tools.py:31) is a compile-time
// constant of its own instantiation: the three sub-steps share one set of constant loads.
// Daily output: 24 warps x 80 registers.  Period means (8 accumulators more per thread): 20 warps x 96 registers.
static cudaError_t launch_lean(const RunArgs &a, cudaStream_t st, int ext)
{
    if (ext == 2)                                // any boundary flags: GENERIC physics, scheme resolved at run time
        return launch_lean_one<SCHEME_RUNTIME, 16, 2, 0, false, 2>(a, st);
    if (ext == 1) {                              // external tsurf / alb rows: daily output only
        if (a.P.n_ksub == 3) return launch_lean_nk<24, 3, 3, false, 1>(a, st);
        return launch_lean_nk<24, 2, 0, false, 1>(a, st);
    }
    if (a.out != nullptr && a.avg_days > 1) {
        if (a.P.n_ksub == 3) return launch_lean_nk<20, 3, 3, true>(a, st);
        return launch_lean_nk<20, 2, 0, true>(a, st);
    }
    if (a.P.n_ksub == 3) {
        const char *w = getenv("SEMIC_WPC");
        const int wpc = w ? atoi(w) : 24;
        if (wpc == 28) return launch_lean_nk<28, 3, 3, false>(a, st);
        if (wpc == 32) return launch_lean_nk<32, 3, 3, false>(a, st);
        return launch_lean_nk<24, 3, 3, false>(a, st);
    }
    if (a.P.n_ksub == 24) {                                                   // BASELINE configs[1]: trip count known, unrolled by 4
        const char *w = getenv("SEMIC_WPC");                                  // experiment knob (tools/ab.py): warps per CTA
        const int wpc = w ? atoi(w) : 24;
        if (wpc == 16) return launch_lean_nk<16, 4, 24, false>(a, st);
        if (wpc == 28) return launch_lean_nk<28, 4, 24, false>(a, st);
        if (wpc == 32) return launch_lean_nk<32, 4, 24, false>(a, st);
        return launch_lean_nk<24, 4, 24, false>(a, st);
    }
    if (a.P.n_ksub >= 8) return launch_lean_nk<24, 4, 0, false>(a, st);
    return launch_lean_nk<24, 2, 0, false>(a, st);
}
#endif
