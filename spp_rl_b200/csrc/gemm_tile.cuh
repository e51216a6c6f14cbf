// FP32 tile GEMM for the fused kernels: C[m,n] = sum_k A(m,k) * B(k,n), CUDA-core FFMA path.  It carries the narrow products
// (heads, ACM, first-layer dW, dX to the inputs), the 64-wide SPP-PPO nets, and -- as the A/B fallback of gemm_umma.cuh -- the
// 256-wide products.  The epilogue functors defined here are shared with the tensor-core path.
//
// Operands live in global memory (L2-resident agent state / activation scratch) and are streamed
// through a 3-stage cp.async (LDGSTS, L2-only) shared-memory pipeline; accumulators stay in
// registers and leave through an epilogue functor (bias+activation store, masked store with column
// sums, Adam+Polyak read-modify-write), so no gradient or pre-activation is ever materialised twice.
//
// Operand layouts (row-major, leading dimension a multiple of 4 floats, pad columns zero):
//   A_KC = true : A is [M x K], contraction contiguous   (forward x, backward dY for dX)
//   A_KC = false: A is [K x M], contraction strided      (dY for dW: K = batch)
//   B is always [K x N] (contraction strided): the forward reads the TRANSPOSED weight copy W^T
//   [in x out], dX reads the natural weight W [out x in], dW reads activations [batch x in].
// Keeping B in outer-product form matters on this register file: with both operands fetched as
// float4-along-k the two multiplicands of every FFMA sit in registers of equal parity (same bank),
// which the first version of this kernel paid for with dispatch stalls on ~40% of its FFMAs
// (profiles/r01_update_v1_summary.md).  With B[k][n..n+3] fragments the b operand's parity follows
// n, the accumulator's parity is free, and the a operand is held in the operand-reuse cache.
// k is accumulated in ascending order with one fmaf per product (the exactness target is 1e-5 relative vs the fp32 reference;
// the tensor-core path needs a three-pass operand split and a per-chunk accumulator drain to meet it, gemm_umma.cuh).
#pragma once
#include "common.cuh"

#ifndef SPP_K4_UNROLL
#define SPP_K4_UNROLL 4      // k-steps of 4 unrolled per loop body (1/2/4/8 measured within 3% of each other; 4 was best)
#endif

namespace spp {

constexpr int kK4Unroll = SPP_K4_UNROLL;
constexpr int TK = 32;                        // contraction chunk per pipeline stage
constexpr int kStages = 3;
constexpr int kPitchKC = TK + 4;              // k-contiguous smem row pitch (conflict-free float4 reads)
constexpr int kOperandFloats = 128 * kPitchKC;          // largest operand tile (128 rows, k-contig)
constexpr int kStageFloats = kOperandFloats + TK * 128; // A region + B region (B is always [TK x TN])
constexpr int kGemmSmemFloats = kStages * kStageFloats; // 26112 floats = 102 KB

template <int TY_, int TX_, int MI_, int NJ_>
struct TileCfg {
    static constexpr int TY = TY_, TX = TX_, MI = MI_, NJ = NJ_;
    static constexpr int TM = TY * MI, TN = TX * NJ;
    static_assert(TY * TX == kThreads, "tile must use the whole CTA");
    static_assert(TM <= 128 && TN <= 128 && MI % 4 == 0 && NJ % 4 == 0, "tile limits");
};
using BigTile = TileCfg<16, 16, 8, 8>;     // 128 x 128, 8x8 per thread
using NarrowTile = TileCfg<32, 8, 4, 4>;   // 128 x 32,  4x4 per thread (heads, ACM, first layers)

// thread -> element mapping.  Columns are always float4 groups: n = 4*tx + j%4 + 4*TX*(j/4).
template <class Cfg, bool A_KC>
__device__ __forceinline__ int row_of(int i, int ty) {
    return A_KC ? ty + Cfg::TY * i : (i / 4) * (4 * Cfg::TY) + 4 * ty + (i % 4);
}
template <class Cfg>
__device__ __forceinline__ int col_of(int j, int tx) { return (j / 4) * (4 * Cfg::TX) + 4 * tx + (j % 4); }

// Per-thread state of one operand's global->shared copies, computed once per tile: source pointer of chunk 0,
// destination offset inside a stage, row validity.  issue() only bumps the pointer and tests the k tail.
template <int ROWS, bool KC>
struct OperandLoader {
    static constexpr int NV = KC ? ROWS * (TK / 4) : TK * (ROWS / 4);
    static constexpr int PER = (NV + kThreads - 1) / kThreads;
    const float* src[PER];
    const float* base;
    int dst[PER], kofs[PER], kstride;
    bool ok[PER];
    __device__ __forceinline__ void init(const float* __restrict__ G, int ld, int R, int r0) {
        base = G;
        kstride = KC ? 1 : ld;
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int c = threadIdx.x + p * kThreads;
            if constexpr (KC) {
                const int row = c / (TK / 4), k4 = c % (TK / 4);
                ok[p] = (c < NV) && (r0 + row < R);
                src[p] = G + (size_t)(r0 + row) * ld + 4 * k4;
                dst[p] = row * kPitchKC + 4 * k4;
                kofs[p] = 4 * k4;
            } else {
                const int kr = c / (ROWS / 4), r4 = c % (ROWS / 4);
                ok[p] = (c < NV) && (r0 + 4 * r4 < R);
                src[p] = G + (size_t)kr * ld + r0 + 4 * r4;
                dst[p] = kr * ROWS + 4 * r4;
                kofs[p] = kr;
            }
        }
    }
    __device__ __forceinline__ void issue(int k0, int K, float* __restrict__ S) const {
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            if (NV % kThreads != 0 && threadIdx.x + p * kThreads >= NV) continue;
            const bool v = ok[p] && (k0 + kofs[p] < K);
            cp_async16(S + dst[p], v ? src[p] + (size_t)k0 * kstride : base, v ? 16 : 0);
        }
    }
};

// One output tile.  Epi interface:
//   template <class Cfg, bool A_KC> void apply(float (&acc)[MI][NJ], int m0, int n0, int M, int N, float* smem);
template <class Cfg, bool A_KC, class Epi>
__device__ __noinline__ void gemm_tile(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                       int ldb, int M, int N, int K, int m0, int n0,
                                       float* __restrict__ smem, const Epi& epi_ref) {
    Epi epi = epi_ref;      // functor fields in registers (see umma_epilogue)
    constexpr int TY = Cfg::TY, TX = Cfg::TX, MI = Cfg::MI, NJ = Cfg::NJ, TM = Cfg::TM, TN = Cfg::TN;
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int nk = (K + TK - 1) / TK;

    float acc[MI][NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    OperandLoader<TM, A_KC> la;
    OperandLoader<TN, false> lb;
    la.init(A, lda, M, m0);
    lb.init(B, ldb, N, n0);
    // prologue
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nk) {
            float* sA = smem + s * kStageFloats;
            la.issue(s * TK, K, sA);
            lb.issue(s * TK, K, sA + kOperandFloats);
        }
        cp_async_commit();
    }

    int slot = 0, pslot = kStages - 1;
    for (int kc = 0; kc < nk; ++kc) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {   // prefetch chunk kc + kStages - 1 into the slot freed by iteration kc - 1
            const int pf = kc + kStages - 1;
            if (pf < nk) {
                float* sP = smem + pslot * kStageFloats;
                la.issue(pf * TK, K, sP);
                lb.issue(pf * TK, K, sP + kOperandFloats);
            }
            cp_async_commit();
            pslot = (pslot + 1 == kStages) ? 0 : pslot + 1;
        }
        const float* sA = smem + slot * kStageFloats;
        const float* sB = sA + kOperandFloats;
        slot = (slot + 1 == kStages) ? 0 : slot + 1;
#pragma unroll kK4Unroll
        for (int k4 = 0; k4 < TK / 4; ++k4) {
            float af[4][MI], bf[4][NJ];
            if constexpr (A_KC) {
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(sA + (ty + TY * i) * kPitchKC + 4 * k4);
                    af[0][i] = v.x; af[1][i] = v.y; af[2][i] = v.z; af[3][i] = v.w;
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int i4 = 0; i4 < MI / 4; ++i4) {
                        const float4 v = *reinterpret_cast<const float4*>(sA + (4 * k4 + kk) * TM + i4 * (4 * TY) + 4 * ty);
                        af[kk][4 * i4 + 0] = v.x; af[kk][4 * i4 + 1] = v.y;
                        af[kk][4 * i4 + 2] = v.z; af[kk][4 * i4 + 3] = v.w;
                    }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int j4 = 0; j4 < NJ / 4; ++j4) {
                    const float4 v = *reinterpret_cast<const float4*>(sB + (4 * k4 + kk) * TN + j4 * (4 * TX) + 4 * tx);
                    bf[kk][4 * j4 + 0] = v.x; bf[kk][4 * j4 + 1] = v.y;
                    bf[kk][4 * j4 + 2] = v.z; bf[kk][4 * j4 + 3] = v.w;
                }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(af[kk][i], bf[kk][j], acc[i][j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // every warp is done reading the pipeline; smem is reusable below and by the next tile
    epi.template apply<Cfg, A_KC>(acc, m0, n0, M, N, smem);
}

// All tiles of one GEMM, distributed round-robin over the CTAs that share an agent.
template <class Cfg, bool A_KC, class Epi>
__device__ __forceinline__ void gemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                     int M, int N, int K, float* __restrict__ smem, Epi epi, int cta_rank = 0,
                                     int cta_count = 1) {
    const int mt = (M + Cfg::TM - 1) / Cfg::TM, nt = (N + Cfg::TN - 1) / Cfg::TN;
    for (int t = cta_rank; t < mt * nt; t += cta_count)
        gemm_tile<Cfg, A_KC, Epi>(A, lda, B, ldb, M, N, K, (t / nt) * Cfg::TM, (t % nt) * Cfg::TN, smem, epi);
}

// L2 prefetch of rows [r0, r0 + nrows) x [c0, c0 + ncols) of a row-major array (ld floats per row), 128-byte lines, all threads.
// The fused kernels' epilogues are latency-bound on reads that miss L2 (every agent's state streams from HBM); prefetching a
// tile's epilogue operands when its main loop starts turns those misses into L2 hits.
__device__ __forceinline__ void prefetch_l2_tile(const float* base, int ld, int r0, int nrows, int c0, int ncols) {
    const int lines_per_row = (ncols + 31) / 32;
    for (int i = threadIdx.x; i < nrows * lines_per_row; i += kThreads) {
        const int r = i / lines_per_row, l = i % lines_per_row;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(r0 + r) * ld + c0 + 32 * l));
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// Streaming variants (L2 evict-first): Adam moments are read and written exactly once per update, 1.5 MB per 256 x 256 layer; with
// 148 agents in flight the 126 MB L2 is better spent on the activations that the next stage re-reads (+1.7 % updates/s measured;
// extending the hint to the weights and Polyak targets changed nothing).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld4_stream(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st4_stream(float* p, const float4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// ---------------------------------------------------------------------------------- epilogues
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };
template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
    if (ACT == ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == ACT_TANH) return tanhf(x);
    return x;
}

// C[m, n] = act(acc + bias[n] (+ t * Sk[m, n])) (* scale[n]); optional pre-scale copy C2 for the backward.
// Pad columns of every row get act(0) * scale = 0 because weight and bias pads are zero.
template <int ACT, bool SCALE, bool ADD>
struct EpiBiasAct {
    float* C; int ldc; const float* bias; const float* scale; float* C2; int ldc2; const float* Sk; int lds; float t;
    // optional (128 x 128 tiles only): the sign pattern of the stored tile, bit i * NJ + j of word [(m0 / 128) * 2 + n0 / 128][thread]
    // = (C[row_of(i), col_of(j)] > 0).  The backward's EpiMaskStore<MASK_RELU> reads these 8 bytes per thread instead of the tile.
    unsigned long long* mask_out = nullptr;
    // optional (128 x 128 tiles of a 256-wide layer, TX = 16): fold a following 256 -> 1 linear head into this epilogue.
    // dot_out[m * 32 + (n0 / 128) * 16 + tx] = sum over the thread's 8 columns of C[m, n] * dot_w[n]; the consumer adds the 32
    // partials of a row in a fixed order.  C may then be null (the activation itself is never stored).
    const float* dot_w = nullptr; float* dot_out = nullptr;
    __device__ __forceinline__ void prefetch(int m0, int rows, int n0, int cols, int M, int N) const {
        if (ADD) prefetch_l2_tile(Sk, lds, m0, min(rows, M - m0), n0, min(cols, N - n0));
    }
    template <class Cfg, bool A_KC>
    __device__ __forceinline__ void apply(float (&acc)[Cfg::MI][Cfg::NJ], int m0, int n0, int M, int N, float*) {
        const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
        constexpr int G4 = Cfg::NJ / 4;
        float4 b[G4], sc[G4], dw[G4];
        unsigned long long bits = 0ull;
#pragma unroll
        for (int g = 0; g < G4; ++g) {
            const int n = n0 + col_of<Cfg>(4 * g, tx);
            b[g] = (n < N) ? ld4(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            dw[g] = (dot_out && n < N) ? ld4(dot_w + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (SCALE) sc[g] = (n < N) ? ld4(scale + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i) {
            const int m = m0 + row_of<Cfg, A_KC>(i, ty);
            if (m >= M) continue;
            float4 ad[G4];
            float dot = 0.f;
            if (ADD) {
#pragma unroll
                for (int g = 0; g < G4; ++g) {
                    const int n = n0 + col_of<Cfg>(4 * g, tx);
                    ad[g] = (n < N) ? ld4(Sk + (size_t)m * lds + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int g = 0; g < G4; ++g) {
                const int n = n0 + col_of<Cfg>(4 * g, tx);
                if (n >= N) continue;
                float4 x = make_float4(acc[i][4 * g] + b[g].x, acc[i][4 * g + 1] + b[g].y, acc[i][4 * g + 2] + b[g].z,
                                       acc[i][4 * g + 3] + b[g].w);
                if (ADD) {
                    x.x = __fadd_rn(x.x, __fmul_rn(t, ad[g].x)); x.y = __fadd_rn(x.y, __fmul_rn(t, ad[g].y));
                    x.z = __fadd_rn(x.z, __fmul_rn(t, ad[g].z)); x.w = __fadd_rn(x.w, __fmul_rn(t, ad[g].w));
                }
                x.x = act_fn<ACT>(x.x); x.y = act_fn<ACT>(x.y); x.z = act_fn<ACT>(x.z); x.w = act_fn<ACT>(x.w);
                if (SCALE) {
                    if (C2) st4(C2 + (size_t)m * ldc2 + n, x);
                    x.x *= sc[g].x; x.y *= sc[g].y; x.z *= sc[g].z; x.w *= sc[g].w;
                }
                if (C) st4(C + (size_t)m * ldc + n, x);
                dot = fmaf(x.x, dw[g].x, dot); dot = fmaf(x.y, dw[g].y, dot); dot = fmaf(x.z, dw[g].z, dot); dot = fmaf(x.w, dw[g].w, dot);
                if constexpr (Cfg::MI * Cfg::NJ == 64)
                    bits |= (unsigned long long)((x.x > 0.f) | ((x.y > 0.f) << 1) | ((x.z > 0.f) << 2) | ((x.w > 0.f) << 3)) << (i * Cfg::NJ + 4 * g);
            }
            if (Cfg::TX == 16 && dot_out) dot_out[(size_t)m * 32 + (n0 / 128) * 16 + tx] = dot;
        }
        if constexpr (Cfg::MI * Cfg::NJ == 64)
            if (mask_out) mask_out[(size_t)((m0 / 128) * 2 + n0 / 128) * kThreads + threadIdx.x] = bits;
    }
};

// Backward through an activation: C[m,n] = acc * act'(H[m,n]); column sums -> bias gradient (atomicAdd).
//   MASK_RELU: act' = (H > 0);  MASK_TANH: act' = 1 - H^2;  MASK_NONE: 1.  ACCUM: C += ...
enum Mask { MASK_NONE = 0, MASK_RELU = 1, MASK_TANH = 2, MASK_RELU_BITS = 3 };      // _BITS: mask_in instead of H (128 x 128 tiles)
template <int MASK, bool COLSUM, bool ACCUM>
struct EpiMaskStore {
    float* C; int ldc; const float* H; int ldh; float* colsum;
    const unsigned long long* mask_in = nullptr;      // MASK_RELU_BITS: the bits EpiBiasAct::mask_out left
    __device__ __forceinline__ void prefetch(int m0, int rows, int n0, int cols, int M, int N) const {
        if (MASK != MASK_NONE) prefetch_l2_tile(H, ldh, m0, min(rows, M - m0), n0, min(cols, N - n0));
        if (ACCUM) prefetch_l2_tile(C, ldc, m0, min(rows, M - m0), n0, min(cols, N - n0));
    }
    template <class Cfg, bool A_KC>
    __device__ __forceinline__ void apply(float (&acc)[Cfg::MI][Cfg::NJ], int m0, int n0, int M, int N, float* smem) {
        const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
        constexpr int G4 = Cfg::NJ / 4;
        float cs[Cfg::NJ];
#pragma unroll
        for (int j = 0; j < Cfg::NJ; ++j) cs[j] = 0.f;
        // issue every global load of the tile first (one latency instead of one per row)
        float4 h[Cfg::MI][G4], old[Cfg::MI][G4];
        static_assert(MASK != MASK_RELU_BITS || Cfg::MI * Cfg::NJ == 64, "bit masks are defined for the 8 x 8 per-thread mapping");
        unsigned long long bits = 0ull;
        if (MASK == MASK_RELU_BITS) bits = mask_in[(size_t)((m0 / 128) * 2 + n0 / 128) * kThreads + threadIdx.x];
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i) {
            const int m = m0 + row_of<Cfg, A_KC>(i, ty);
#pragma unroll
            for (int g = 0; g < G4; ++g) {
                const int n = n0 + col_of<Cfg>(4 * g, tx);
                h[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
                old[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MASK != MASK_NONE && MASK != MASK_RELU_BITS && m < M && n < N) h[i][g] = ld4(H + (size_t)m * ldh + n);
                if (ACCUM && m < M && n < N) old[i][g] = ld4(C + (size_t)m * ldc + n);
            }
        }
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i) {
            const int m = m0 + row_of<Cfg, A_KC>(i, ty);
            if (m >= M) continue;
#pragma unroll
            for (int g = 0; g < G4; ++g) {
                const int n = n0 + col_of<Cfg>(4 * g, tx);
                if (n >= N) continue;
                float x[4] = {acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]};
                const float hh[4] = {h[i][g].x, h[i][g].y, h[i][g].z, h[i][g].w};
                const float oo[4] = {old[i][g].x, old[i][g].y, old[i][g].z, old[i][g].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (MASK == MASK_RELU) x[e] = (hh[e] > 0.f) ? x[e] : 0.f;
                    if (MASK == MASK_RELU_BITS) x[e] = ((bits >> (i * Cfg::NJ + 4 * g + e)) & 1ull) ? x[e] : 0.f;
                    if (MASK == MASK_TANH) x[e] = x[e] * (1.f - hh[e] * hh[e]);
                    if (ACCUM) x[e] += oo[e];
                    if (COLSUM) cs[4 * g + e] += x[e];
                }
                st4(C + (size_t)m * ldc + n, make_float4(x[0], x[1], x[2], x[3]));
            }
        }
        if constexpr (COLSUM) {
            float* red = smem;   // [TY][TN], pipeline is drained
#pragma unroll
            for (int j = 0; j < Cfg::NJ; ++j) red[ty * Cfg::TN + col_of<Cfg>(j, tx)] = cs[j];
            __syncthreads();
            for (int c = threadIdx.x; c < Cfg::TN; c += kThreads) {
                float s = 0.f;
#pragma unroll 4
                for (int t = 0; t < Cfg::TY; ++t) s += red[t * Cfg::TN + c];
                if (n0 + c < N) atomicAdd(colsum + n0 + c, s);
            }
            __syncthreads();
        }
    }
};

// torch.optim.Adam (single-tensor CPU form) on one element.
struct AdamScalars {
    float lr_over_bc1;     // step_size = lr / (1 - beta1^t)
    float bc2_sqrt;        // sqrt(1 - beta2^t)
};
// sqrt and the two divisions use the MUFU approximations (<= 2 ulp): the update term lr*m/(sqrt(v)+eps) then carries
// ~2e-7 relative error, far inside the 1e-5 parity budget, at a sixth of the instruction count of IEEE div/sqrt.
__device__ __forceinline__ float adam_element(float w, float g, float& m, float& v, const AdamScalars& s) {
    m = fmaf(0.1f, __fsub_rn(g, m), m);                                   // lerp_(g, 1 - beta1): ATen's vector path is an fma
    v = __fadd_rn(__fmul_rn(v, 0.999f), __fmul_rn(__fmul_rn(0.001f, g), g));   // mul_(beta2).addcmul_(g, g, 1 - beta2)
    float sq;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
    const float denom = __fadd_rn(__fdividef(sq, s.bc2_sqrt), 1e-8f);          // (sqrt(v) / sqrt(bc2)).add_(eps)
    return __fadd_rn(w, __fdividef(__fmul_rn(-s.lr_over_bc1, m), denom));       // addcdiv_: w + (value * m) / denom
}

// dW tile (TN form, thread owns 4x4 blocks) -> Adam.  Element (m, n) of the tile lives at
//   R[m * ldr + n]  in the "row" copy (moments Mo/Vo use this layout) and at
//   Cc[n * ldcc + m] in the "column" copy (may be null), with optional Polyak target in column layout:
//   Tc[n * ldcc + m] = Tc * (1 - tau) + tau * w_new   (SAC.update_target_q / DDPG.update_target_nets);
//   Tr is the same blend for a target kept in the row layout.
// For ordinary layers R = natural W [out x in] and Cc = W^T [in x out] (what the forward reads);
// for the actor heads the tile is computed transposed, so R = W^T and Cc = W.
struct EpiAdam {
    float* R; float* Mo; float* Vo; int ldr; float* Cc; float* Tc; float* Tr; int ldcc; AdamScalars s; float tau, one_minus_tau;
    __device__ __forceinline__ void prefetch(int m0, int rows, int n0, int cols, int M, int N) const {
        const int nr = min(rows, M - m0), nc = min(cols, N - n0);
        prefetch_l2_tile(R, ldr, m0, nr, n0, nc);
        prefetch_l2_tile(Mo, ldr, m0, nr, n0, nc);
        prefetch_l2_tile(Vo, ldr, m0, nr, n0, nc);
        if (Tr) prefetch_l2_tile(Tr, ldr, m0, nr, n0, nc);
        if (Tc) prefetch_l2_tile(Tc, ldcc, n0, nc, m0, nr);      // column copy: element (m, n) at Tc[n * ldcc + m]
    }
    template <class Cfg, bool A_KC>
    __device__ __forceinline__ void apply(float (&acc)[Cfg::MI][Cfg::NJ], int m0, int n0, int M, int N, float*) {
        static_assert(!A_KC, "Adam epilogue expects the dW (TN) mapping");
        const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
        const uint64_t pol = l2_evict_first_policy();
#pragma unroll
        for (int i4 = 0; i4 < Cfg::MI / 4; ++i4) {
            const int mb = m0 + row_of<Cfg, false>(4 * i4, ty);
            if (mb >= M) continue;
#pragma unroll
            for (int j4 = 0; j4 < Cfg::NJ / 4; ++j4) {
                const int nb = n0 + col_of<Cfg>(4 * j4, tx);
                if (nb >= N) continue;
                float4 w[4], mo[4], vo[4], tg[4];      // tg: the Polyak target block, row layout (Tr) or column layout (Tc) -- never both
#pragma unroll
                for (int r = 0; r < 4; ++r) {   // rows mb..mb+3; every global load of the block is issued up front
                    const size_t o = (size_t)(mb + r) * ldr + nb;
                    const bool ok = (mb + r) < M;
                    w[r] = ok ? ld4(R + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                    mo[r] = ok ? ld4_stream(Mo + o, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
                    vo[r] = ok ? ld4_stream(Vo + o, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
                    tg[r] = Tr ? (ok ? ld4(Tr + o) : make_float4(0.f, 0.f, 0.f, 0.f))
                               : ((Tc && nb + r < N) ? ld4(Tc + (size_t)(nb + r) * ldcc + mb) : make_float4(0.f, 0.f, 0.f, 0.f));
                }
                float wn[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float* wm = reinterpret_cast<float*>(&mo[r]);
                    float* wv = reinterpret_cast<float*>(&vo[r]);
                    const float* ww = reinterpret_cast<const float*>(&w[r]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) wn[r][e] = adam_element(ww[e], acc[4 * i4 + r][4 * j4 + e], wm[e], wv[e], s);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if ((mb + r) >= M) continue;
                    const size_t o = (size_t)(mb + r) * ldr + nb;
                    const float4 nv = make_float4(wn[r][0], wn[r][1], wn[r][2], wn[r][3]);
                    st4(R + o, nv);
                    st4_stream(Mo + o, mo[r], pol);
                    st4_stream(Vo + o, vo[r], pol);
                    if (Tr) {   // Polyak target kept in the row layout (actor heads of DDPG)
                        const float4 t = tg[r];
                        st4(Tr + o, make_float4(__fadd_rn(__fmul_rn(t.x, one_minus_tau), __fmul_rn(tau, nv.x)),
                                                __fadd_rn(__fmul_rn(t.y, one_minus_tau), __fmul_rn(tau, nv.y)),
                                                __fadd_rn(__fmul_rn(t.z, one_minus_tau), __fmul_rn(tau, nv.z)),
                                                __fadd_rn(__fmul_rn(t.w, one_minus_tau), __fmul_rn(tau, nv.w))));
                    }
                }
                if (Cc) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {   // column copy: 4 consecutive m for column nb + e
                        if (nb + e >= N) continue;
                        const size_t o = (size_t)(nb + e) * ldcc + mb;
                        const float4 nv = make_float4(wn[0][e], wn[1][e], wn[2][e], wn[3][e]);
                        st4(Cc + o, nv);
                        if (Tc && !Tr) {
                            const float4 t = tg[e];
                            st4(Tc + o, make_float4(__fadd_rn(__fmul_rn(t.x, one_minus_tau), __fmul_rn(tau, nv.x)),
                                                    __fadd_rn(__fmul_rn(t.y, one_minus_tau), __fmul_rn(tau, nv.y)),
                                                    __fadd_rn(__fmul_rn(t.z, one_minus_tau), __fmul_rn(tau, nv.z)),
                                                    __fadd_rn(__fmul_rn(t.w, one_minus_tau), __fmul_rn(tau, nv.w))));
                        }
                    }
                }
            }
        }
    }
};

}  // namespace spp
