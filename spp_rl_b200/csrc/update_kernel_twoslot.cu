// The fused update burst kernels once more, built with the two-slot tcgen05 main loop (umma_mainloop, gemm_umma.cuh) instead of the
// landing-zone loop: the reduced-precision variant (spp_set_gemm_path(2): one tf32 pass, whole K in one TMEM accumulator) is served by
// this build -- its single-pass code path is the one every round-2 measurement and tolerance statement of the variant was made on --
// and `SPP_UMMA_LOOP=twoslot` selects it for the default path as well (A/B of the two loops inside one library: bitwise equal results).
// Same source, other kernel / launcher names (one kernel body keeps ONE main loop in its call graph, see gemm256_umma).
#define SPP_NO_LANDING_ZONE
#define update_burst_kernel update_burst_twoslot_kernel
#define update_burst_interleaved_kernel update_burst_interleaved_twoslot_kernel
#define launch_update_burst launch_update_burst_twoslot
#include "update_kernel.cu"
