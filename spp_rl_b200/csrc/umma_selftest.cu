// Self-test of the tcgen05 building blocks (umma.cuh): C[128 x 128] = A * B for one CTA, every operand-layout combination
// the fused kernels need, in 1-pass tf32 or 3-pass split (fp32-accurate) mode.  Driven from tests through the C ABI.
//   a_mn = 0: A is [128 x K] row-major (k contiguous)      a_mn = 1: A is [K x 128] row-major (m contiguous)
//   b_mn = 0: B is [128 x K] row-major (k contiguous)      b_mn = 1: B is [K x 128] row-major (n contiguous)
#include <string>

#include "../../include/spp_rl_b200.h"
#include "common.cuh"
#include "gemm_umma.cuh"

int spp_set_error_(int code, const std::string& msg);
void spp_count_launch_();

namespace spp {

using namespace umma;

template <bool MN>
__device__ __forceinline__ void fill_tile(const float* __restrict__ G, int ld, int k0, int K, unsigned char* tile_hi,
                                          unsigned char* tile_lo, bool split) {
    // 128 x 32 (K-major) or 32 x 128 (MN-major) floats = 1024 16-byte chunks, 4 per thread
    for (int c = threadIdx.x; c < 1024; c += kThreads) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t off;
        if (!MN) {
            const int row = c >> 3, chunk = c & 7;
            if (k0 + chunk * 4 < K) v = *reinterpret_cast<const float4*>(G + (size_t)row * ld + k0 + chunk * 4);
            off = kmajor_offset(row, chunk);
        } else {
            const int k = c >> 5, chunk = c & 31;
            if (k0 + k < K) v = *reinterpret_cast<const float4*>(G + (size_t)(k0 + k) * ld + chunk * 4);
            off = mnmajor_offset(k, chunk);
        }
        if (split) {
            float4 h, l;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(tile_hi + off) = h;
            *reinterpret_cast<float4*>(tile_lo + off) = l;
        } else {
            *reinterpret_cast<float4*>(tile_hi + off) = v;
        }
    }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1) umma_selftest_kernel(const float* A, int lda, const float* B, int ldb, float* C, int K,
                                                                   int split) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* a_hi = smem; unsigned char* a_lo = smem + kTileBytes;
    unsigned char* b_hi = smem + 2 * kTileBytes; unsigned char* b_lo = smem + 3 * kTileBytes;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    if (warp_id() == 0) tmem_alloc<128>(&tmem_base_s);
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(128, 128, A_MN ? 1 : 0, B_MN ? 1 : 0);
    uint32_t phase = 0;
    const int nk = (K + kChunkK - 1) / kChunkK;
    for (int kc = 0; kc < nk; ++kc) {
        fill_tile<A_MN>(A, lda, kc * kChunkK, K, a_hi, a_lo, split != 0);
        fill_tile<B_MN>(B, ldb, kc * kChunkK, K, b_hi, b_lo, split != 0);
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t dah = A_MN ? mnmajor_desc(smem_u32(a_hi), ks) : kmajor_desc(smem_u32(a_hi), ks);
                const uint64_t dbh = B_MN ? mnmajor_desc(smem_u32(b_hi), ks) : kmajor_desc(smem_u32(b_hi), ks);
                mma_tf32(tmem_d, dah, dbh, idesc, (kc | ks) ? 1u : 0u);
                if (split) {
                    const uint64_t dal = A_MN ? mnmajor_desc(smem_u32(a_lo), ks) : kmajor_desc(smem_u32(a_lo), ks);
                    const uint64_t dbl = B_MN ? mnmajor_desc(smem_u32(b_lo), ks) : kmajor_desc(smem_u32(b_lo), ks);
                    mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                    mma_tf32(tmem_d, dal, dbh, idesc, 1u);
                }
            }
            commit(&mbar);
        }
        mbar_wait(&mbar, phase);      // single-buffered: wait until the tensor core has consumed this chunk
        phase ^= 1;
        __syncthreads();
    }
    fence_after_sync();
    // epilogue: warp w reads TMEM lanes 32*(w%4).., columns 64*(w/4)..
    const int row = 32 * (warp_id() & 3) + lane_id();
    const int col0 = 64 * (warp_id() >> 2);
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * (warp_id() & 3)) << 16) + col0 + cb * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) C[(size_t)row * 128 + col0 + cb * 16 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp_id() == 0) tmem_dealloc<128>(tmem_d);
}

// the full pipelined GEMM of the fused kernels (gemm_umma.cuh) with a plain store epilogue
template <bool A_KM, bool B_KM>
__global__ void __launch_bounds__(kThreads, 1) umma_gemm_selftest_kernel(const float* A, int lda, const float* B, int ldb, float* C, int M,
                                                                        int K, int reps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t mbar[kUmmaSlots];
    __shared__ uint32_t tmem_base_s;
    if (warp_id() == 0) tmem_alloc<kUmmaTmemCols>(&tmem_base_s);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kUmmaSlots; ++s) mbar_init(mbar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    __shared__ UmmaCtx u;      // pipeline state is shared: thread 0 publishes the barrier parities after every tile
    if (threadIdx.x == 0) u = UmmaCtx{smem, mbar, tmem_base_s, 0u, (uint32_t)(reps >> 16)};
    reps &= 0xFFFF;
    __syncthreads();
    EpiMaskStore<MASK_NONE, false, false> epi{C, 256, nullptr, 0, nullptr};
    for (int r = 0; r < reps; ++r) gemm256_umma<A_KM, B_KM>(A, lda, B, ldb, M, K, u, epi);
    fence_before_sync();
    __syncthreads();
    if (warp_id() == 0) tmem_dealloc<kUmmaTmemCols>(u.tmem);
}


// Hardware probes for the SPP-PPO critic kernel (ppo_critic_tc.cu), single pass, K = 32:
//   mode 1: the A operand K-major but stored with the SWIZZLE_128B_BASE32B pattern (the byte image of an MN-major tile): can ONE
//           shared-memory image of a [rows x 32] block serve as K-major A (forward) and as MN-major operand (dW products)?
//   mode 2: M = 64 (A rows 0..63): where do the 64 accumulator rows land in TMEM?  All 128 lanes x 128 columns are dumped.
__global__ void __launch_bounds__(kThreads, 1) umma_probe_kernel(const float* A, const float* B, float* C, int mode) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_t = smem; unsigned char* b_t = smem + kTileBytes;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    if (warp_id() == 0) tmem_alloc<128>(&tmem_base_s);
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    for (int c = threadIdx.x; c < 1024; c += kThreads) {      // 128 rows x 8 chunks of 16 B
        const int row = c >> 3, chunk = c & 7;
        const float4 va = *reinterpret_cast<const float4*>(A + (size_t)row * 32 + chunk * 4);
        const float4 vb = *reinterpret_cast<const float4*>(B + (size_t)row * 32 + chunk * 4);
        uint32_t off_a = kmajor_offset(row, chunk);
        if (mode == 1) off_a = (uint32_t)(row * 128 + ((((chunk >> 1) ^ (row & 3)) << 5) | ((chunk & 1) << 4)));
        if (mode == 3) off_a = (uint32_t)(((row >> 3) * 8 + chunk) * 128 + (row & 7) * 16);      // no swizzle: core matrices of 8 rows x 16 B
        *reinterpret_cast<float4*>(a_t + off_a) = va;
        if (mode < 4) *reinterpret_cast<float4*>(b_t + kmajor_offset(row, chunk)) = vb;
    }
    if (mode >= 4) {      // B given as [32 k-rows][128 n]: no-swizzle image, core matrix = 8 k-rows x 16 B of n
        for (int c = threadIdx.x; c < 1024; c += kThreads) {
            const int k = c >> 5, chunk = c & 31;
            *reinterpret_cast<float4*>(b_t + ((k >> 3) * 32 + chunk) * 128 + (k & 7) * 16) = *reinterpret_cast<const float4*>(B + (size_t)k * 128 + chunk * 4);
        }
    }
    // clear the accumulator lanes first so that untouched lanes read as a marker
    {
        const uint32_t t = tmem_d + ((uint32_t)(32 * (warp_id() & 3)) << 16) + 64 * (warp_id() >> 2);
        for (int cb = 0; cb < 4; ++cb)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(t + cb * 16), "r"(0x7fc00000u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_tf32(mode == 2 ? 64 : 128, 128, 0, mode >= 4 ? 1 : 0);
        for (int ks = 0; ks < 4; ++ks) {
            uint64_t da = kmajor_desc(smem_u32(a_t), ks), db = kmajor_desc(smem_u32(b_t), ks);
            if (mode == 1) da = make_desc(smem_u32(a_t) + ks * 32, 16, 1024, 1);
            if (mode == 3) da = make_desc(smem_u32(a_t) + ks * 256, 128, 1024, 0);
            if (mode == 4) db = make_desc(smem_u32(b_t) + ks * 4096, 128, 4096, 0);      // LBO along n, SBO along k
            if (mode == 5) db = make_desc(smem_u32(b_t) + ks * 4096, 4096, 128, 0);      // the other way round
            mma_tf32(tmem_d, da, db, idesc, ks ? 1u : 0u);
        }
        commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    const int row = 32 * (warp_id() & 3) + lane_id();
    const int col0 = 64 * (warp_id() >> 2);
    for (int cb = 0; cb < 4; ++cb) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * (warp_id() & 3)) << 16) + col0 + cb * 16, v);
        for (int i = 0; i < 16; ++i) C[(size_t)row * 128 + col0 + cb * 16 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp_id() == 0) tmem_dealloc<128>(tmem_d);
}

}  // namespace spp

extern "C" int spp_umma_probe(int mode, const float* A, const float* B, float* C) {      // A, B [128][32]; C [128][128] (TMEM lanes x columns)
    using namespace spp;
    if (!A || !B || !C || mode < 0 || mode > 5) return spp_set_error_(SPP_ERR_ARG, "spp_umma_probe: bad argument");
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    cudaError_t e;
#define UCK(x) if ((e = (x)) != cudaSuccess) { cudaFree(dA); cudaFree(dB); cudaFree(dC); return spp_set_error_(SPP_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e)); }
    UCK(cudaMalloc(&dA, 128 * 32 * 4)); UCK(cudaMalloc(&dB, 128 * 32 * 4)); UCK(cudaMalloc(&dC, 128 * 128 * 4));
    UCK(cudaMemcpy(dA, A, 128 * 32 * 4, cudaMemcpyHostToDevice)); UCK(cudaMemcpy(dB, B, 128 * 32 * 4, cudaMemcpyHostToDevice));
    const size_t smem = 2 * umma::kTileBytes + 1024;
    UCK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_kernel<<<1, kThreads, smem>>>(dA, dB, dC, mode);
    UCK(cudaGetLastError());
    spp_count_launch_();
    UCK(cudaDeviceSynchronize());
    UCK(cudaMemcpy(C, dC, 128 * 128 * 4, cudaMemcpyDeviceToHost));
#undef UCK
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return SPP_OK;
}

// C[M x 256] = A . B through gemm256_umma; a_km: A stored [M][K] (1) or [K][M] (0); b_km: B stored [256][K] (1) or [K][256] (0).
// M <= 256 and a multiple of 4, K a multiple of 4.  reps > 1 repeats the GEMM (pipeline state carried over); ms_out (optional)
// receives the kernel time.
extern "C" int spp_umma_gemm_selftest(int a_km, int b_km, int M, int K, int reps, const float* A, const float* B, float* C, float* ms_out) {
    using namespace spp;
    if (!A || !B || !C || M < 1 || M > 256 || K < 1 || (K % 4) != 0 || (M % 4) != 0 || reps < 1)
        return spp_set_error_(SPP_ERR_ARG, "spp_umma_gemm_selftest: bad argument");
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e;
#define UCK(x) if ((e = (x)) != cudaSuccess) { cudaFree(dA); cudaFree(dB); cudaFree(dC); return spp_set_error_(SPP_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e)); }
    UCK(cudaMalloc(&dA, (size_t)M * K * 4)); UCK(cudaMalloc(&dB, (size_t)256 * K * 4)); UCK(cudaMalloc(&dC, (size_t)M * 256 * 4));
    UCK(cudaMemcpy(dA, A, (size_t)M * K * 4, cudaMemcpyHostToDevice)); UCK(cudaMemcpy(dB, B, (size_t)256 * K * 4, cudaMemcpyHostToDevice));
    UCK(cudaMemset(dC, 0, (size_t)M * 256 * 4));
    const int lda = a_km ? K : M, ldb = b_km ? K : 256;
    const size_t smem = kUmmaSmemBytes + 1024;
    UCK(cudaEventCreate(&e0)); UCK(cudaEventCreate(&e1));
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (r != cudaSuccess) return r;
        cudaEventRecord(e0);
        kern<<<1, kThreads, smem>>>(dA, lda, dB, ldb, dC, M, K, reps);
        cudaEventRecord(e1);
        return cudaGetLastError();
    };
    if (a_km && b_km) { UCK(launch(umma_gemm_selftest_kernel<true, true>)); }
    else if (a_km && !b_km) { UCK(launch(umma_gemm_selftest_kernel<true, false>)); }
    else if (!a_km && b_km) { UCK(launch(umma_gemm_selftest_kernel<false, true>)); }
    else { UCK(launch(umma_gemm_selftest_kernel<false, false>)); }
    spp_count_launch_();
    UCK(cudaDeviceSynchronize());
    float ms = 0.f;
    UCK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_out) *ms_out = ms;
    UCK(cudaMemcpy(C, dC, (size_t)M * 256 * 4, cudaMemcpyDeviceToHost));
#undef UCK
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return SPP_OK;
}

extern "C" int spp_umma_selftest(int a_mn, int b_mn, int K, int split, const float* A, const float* B, float* C) {
    using namespace spp;
    if (!A || !B || !C || K < 1 || (K % 4) != 0) return spp_set_error_(SPP_ERR_ARG, "spp_umma_selftest: bad argument (K must be a multiple of 4)");
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    const size_t na = (size_t)128 * K;
    cudaError_t e;
#define UCK(x) if ((e = (x)) != cudaSuccess) { cudaFree(dA); cudaFree(dB); cudaFree(dC); return spp_set_error_(SPP_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e)); }
    UCK(cudaMalloc(&dA, na * 4)); UCK(cudaMalloc(&dB, na * 4)); UCK(cudaMalloc(&dC, 128 * 128 * 4));
    UCK(cudaMemcpy(dA, A, na * 4, cudaMemcpyHostToDevice)); UCK(cudaMemcpy(dB, B, na * 4, cudaMemcpyHostToDevice));
    const int lda = a_mn ? 128 : K, ldb = b_mn ? 128 : K;
    const size_t smem = 4 * umma::kTileBytes + 1024;
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (r != cudaSuccess) return r;
        kern<<<1, kThreads, smem>>>(dA, lda, dB, ldb, dC, K, split);
        return cudaGetLastError();
    };
    if (!a_mn && !b_mn) { UCK(launch(umma_selftest_kernel<false, false>)); }
    else if (!a_mn && b_mn) { UCK(launch(umma_selftest_kernel<false, true>)); }
    else if (a_mn && !b_mn) { UCK(launch(umma_selftest_kernel<true, false>)); }
    else { UCK(launch(umma_selftest_kernel<true, true>)); }
    spp_count_launch_();
    UCK(cudaDeviceSynchronize());
    UCK(cudaMemcpy(C, dC, 128 * 128 * 4, cudaMemcpyDeviceToHost));
#undef UCK
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return SPP_OK;
}
