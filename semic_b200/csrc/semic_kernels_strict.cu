// Strict arithmetic instantiation: reference operation order, IEEE division, no FMA contraction
// (this file is compiled with -fmad=false).  Used for parity debugging and the flip report.
#define SEMIC_STRICT 1
#define SEMIC_MODE_NS strict_mode
#include "semic_kernels.cuh"

namespace semic {
namespace strict_mode {
#include "semic_launch.inl"
}  // namespace strict_mode

cudaError_t launch_run_strict(const RunArgs &a, bool generic, cudaStream_t st, int *launches)
{
    if (launches) ++*launches;
    return generic ? strict_mode::launch_scheme<true, 1, 8>(a, st) : strict_mode::launch_scheme<false, 1, 8>(a, st);
}
}  // namespace semic
