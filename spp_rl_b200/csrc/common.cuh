// Device helpers shared by the SPP-RL kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spp {

constexpr int kThreads = 256;          // one CTA = 8 warps, one CTA per SM (persistent)
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// 16-byte async copy global->shared (LDGSTS), L2-only (.cg); src_bytes = 0 zero-fills.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in a fixed order (deterministic): warp tree, then warp 0 adds the 8 partials.
// `red` points at >= kWarps floats of shared memory.  Result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[warp_id()] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[w];
    return t;
}

// ---------------------------------------------------------------- Philox4x32-10 (counter-based RNG)
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __device__ static __forceinline__ uint4 round4(uint4 c, uint2 k) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    }
    __device__ static __forceinline__ uint4 gen(uint64_t seed, uint64_t stream, uint64_t ctr) {
        uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        uint4 c = make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream, (uint32_t)(stream >> 32));
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            c = round4(c, k);
            k.x += W0;
            k.y += W1;
        }
        return c;
    }
};

// One standard normal from two 32-bit words (Box-Muller, cosine branch).
__device__ __forceinline__ float normal_from_bits(uint32_t a, uint32_t b) {
    float u1 = (float)(a >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);   // (0,1)
    float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// torch.nn.functional.softplus(x) with beta = 1, threshold = 20.
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

}  // namespace spp
