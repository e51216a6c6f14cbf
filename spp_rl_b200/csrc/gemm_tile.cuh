// FP32 tile GEMM for the fused update kernels: C[m,n] = sum_k A(m,k) * B(n,k), CUDA-core FFMA path.
//
// Operands live in global memory (L2-resident agent state / activation scratch) and are streamed
// through a 4-stage cp.async (LDGSTS, L2-only) shared-memory pipeline; accumulators stay in
// registers and leave through an epilogue functor (bias+activation store, masked store with column
// sums, Adam+Polyak read-modify-write), so no gradient or pre-activation is ever materialised twice.
//
// Operand layouts (row-major, leading dimension a multiple of 4 floats, pad columns zero):
//   A_KC = true : A is [M x K], contraction contiguous   (forward x, backward dY for dX)
//   A_KC = false: A is [K x M], contraction strided      (dY for dW: K = batch)
//   B_KC = true : B is [N x K]                            (torch Linear weight in forward)
//   B_KC = false: B is [K x N]                            (weight in dX; activations in dW)
// The exactness target (1e-5 relative vs the fp32 reference) is why this path is FFMA and not
// single-pass bf16/tf32 tensor cores; k is accumulated in ascending order.
#pragma once
#include "common.cuh"

namespace spp {

constexpr int TK = 16;                        // contraction chunk per pipeline stage
constexpr int kStages = 4;
constexpr int kPitchKC = TK + 4;              // k-contiguous smem row pitch (conflict-free float4 reads)
constexpr int kOperandFloats = 128 * kPitchKC;          // largest operand tile (128 rows, k-contig)
constexpr int kStageFloats = 2 * kOperandFloats;        // A region + B region
constexpr int kGemmSmemFloats = kStages * kStageFloats; // 20480 floats = 80 KB

template <int TY_, int TX_, int MI_, int NJ_>
struct TileCfg {
    static constexpr int TY = TY_, TX = TX_, MI = MI_, NJ = NJ_;
    static constexpr int TM = TY * MI, TN = TX * NJ;
    static_assert(TY * TX == kThreads, "tile must use the whole CTA");
    static_assert(TM <= 128 && TN <= 128 && MI % 4 == 0 && NJ % 4 == 0, "tile limits");
};
using BigTile = TileCfg<16, 16, 8, 8>;     // 128 x 128, 8x8 per thread
using NarrowTile = TileCfg<32, 8, 4, 4>;   // 128 x 32,  4x4 per thread (heads, ACM, first layers)

template <int ROWS, bool KC>
__device__ __forceinline__ void load_operand(const float* __restrict__ G, int ld, int R, int K, int r0,
                                             int k0, float* __restrict__ S) {
    const int tid = threadIdx.x;
    if constexpr (KC) {
        constexpr int NV = ROWS * (TK / 4);
#pragma unroll
        for (int c = tid; c < NV; c += kThreads) {
            const int row = c / (TK / 4), k4 = c % (TK / 4);
            const int gr = r0 + row, gk = k0 + 4 * k4;
            const bool ok = (gr < R) && (gk < K);
            const float* src = ok ? G + (size_t)gr * ld + gk : G;
            cp_async16(S + row * kPitchKC + 4 * k4, src, ok ? 16 : 0);
        }
    } else {
        constexpr int NV = TK * (ROWS / 4);
#pragma unroll
        for (int c = tid; c < NV; c += kThreads) {
            const int kr = c / (ROWS / 4), r4 = c % (ROWS / 4);
            const int gk = k0 + kr, gr = r0 + 4 * r4;
            const bool ok = (gk < K) && (gr < R);
            const float* src = ok ? G + (size_t)gk * ld + gr : G;
            cp_async16(S + kr * ROWS + 4 * r4, src, ok ? 16 : 0);
        }
    }
}

// One output tile.  Epi interface:
//   static constexpr bool kColSum;            // reduce the returned values over m, atomicAdd into colsum_dst()
//   float* colsum_dst();
//   template<int NJ> void row(int m, const int (&n)[NJ], float (&v)[NJ], int N);  // v in: acc, out: value for colsum
template <class Cfg, bool A_KC, bool B_KC, class Epi>
__device__ __noinline__ void gemm_tile(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                          int ldb, int M, int N, int K, int m0, int n0,
                                          float* __restrict__ smem, Epi& epi) {
    constexpr int TY = Cfg::TY, TX = Cfg::TX, MI = Cfg::MI, NJ = Cfg::NJ, TM = Cfg::TM, TN = Cfg::TN;
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int nk = (K + TK - 1) / TK;

    float acc[MI][NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    // prologue
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nk) {
            float* sA = smem + s * kStageFloats;
            load_operand<TM, A_KC>(A, lda, M, K, m0, s * TK, sA);
            load_operand<TN, B_KC>(B, ldb, N, K, n0, s * TK, sA + kOperandFloats);
        }
        cp_async_commit();
    }

    for (int kc = 0; kc < nk; ++kc) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {   // prefetch chunk kc + kStages - 1 into the slot freed by iteration kc - 1
            const int pf = kc + kStages - 1;
            if (pf < nk) {
                float* sA = smem + (pf % kStages) * kStageFloats;
                load_operand<TM, A_KC>(A, lda, M, K, m0, pf * TK, sA);
                load_operand<TN, B_KC>(B, ldb, N, K, n0, pf * TK, sA + kOperandFloats);
            }
            cp_async_commit();
        }
        const float* sA = smem + (kc % kStages) * kStageFloats;
        const float* sB = sA + kOperandFloats;
#pragma unroll
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
            if constexpr (B_KC) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(sB + (tx + TX * j) * kPitchKC + 4 * k4);
                    bf[0][j] = v.x; bf[1][j] = v.y; bf[2][j] = v.z; bf[3][j] = v.w;
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int j4 = 0; j4 < NJ / 4; ++j4) {
                        const float4 v = *reinterpret_cast<const float4*>(sB + (4 * k4 + kk) * TN + j4 * (4 * TX) + 4 * tx);
                        bf[kk][4 * j4 + 0] = v.x; bf[kk][4 * j4 + 1] = v.y;
                        bf[kk][4 * j4 + 2] = v.z; bf[kk][4 * j4 + 3] = v.w;
                    }
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

    int ncol[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
        ncol[j] = n0 + (B_KC ? tx + TX * j : (j / 4) * (4 * TX) + 4 * tx + (j % 4));
    float cs[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) cs[j] = 0.f;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = m0 + (A_KC ? ty + TY * i : (i / 4) * (4 * TY) + 4 * ty + (i % 4));
        if (m < M) {
            epi.template row<NJ>(m, ncol, acc[i], N);
            if constexpr (Epi::kColSum) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) cs[j] += acc[i][j];
            }
        }
    }
    if constexpr (Epi::kColSum) {
        float* red = smem;   // [TY][TN]
#pragma unroll
        for (int j = 0; j < NJ; ++j) red[ty * TN + (ncol[j] - n0)] = cs[j];
        __syncthreads();
        for (int c = tid; c < TN; c += kThreads) {
            float s = 0.f;
#pragma unroll 4
            for (int t = 0; t < TY; ++t) s += red[t * TN + c];
            if (n0 + c < N) atomicAdd(epi.colsum_dst() + n0 + c, s);
        }
        __syncthreads();
    }
}

// All tiles of one GEMM, distributed round-robin over the CTAs that share an agent.
template <class Cfg, bool A_KC, bool B_KC, class Epi>
__device__ __forceinline__ void gemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                     int M, int N, int K, float* __restrict__ smem, Epi epi, int cta_rank = 0,
                                     int cta_count = 1) {
    const int mt = (M + Cfg::TM - 1) / Cfg::TM, nt = (N + Cfg::TN - 1) / Cfg::TN;
    for (int t = cta_rank; t < mt * nt; t += cta_count)
        gemm_tile<Cfg, A_KC, B_KC, Epi>(A, lda, B, ldb, M, N, K, (t / nt) * Cfg::TM, (t % nt) * Cfg::TN, smem, epi);
}

// ---------------------------------------------------------------------------------- epilogues
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

// C[m, n] = act(acc + bias[n]) (* scale[n]); optional second copy (pre-scale) for backward.
template <int ACT, bool SCALE>
struct EpiBiasAct {
    static constexpr bool kColSum = false;
    float* C; int ldc; const float* bias; const float* scale; float* C2; int ldc2;
    __device__ float* colsum_dst() { return nullptr; }
    template <int NJ>
    __device__ __forceinline__ void row(int m, const int (&n)[NJ], float (&v)[NJ], int N) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (n[j] < N) {
                float x = v[j] + bias[n[j]];
                if (ACT == ACT_RELU) x = fmaxf(x, 0.f);
                if (ACT == ACT_TANH) x = tanhf(x);
                if (SCALE) {
                    if (C2) C2[(size_t)m * ldc2 + n[j]] = x;
                    x = x * scale[n[j]];
                }
                C[(size_t)m * ldc + n[j]] = x;
            }
        }
    }
};

// C[m, n] = tanh(acc + bias[n] + t * S[m, n])   (BasicAcM hidden layer with its skip connection)
struct EpiBiasAddAct {
    static constexpr bool kColSum = false;
    float* C; int ldc; const float* bias; const float* Sk; int lds; float t;
    __device__ float* colsum_dst() { return nullptr; }
    template <int NJ>
    __device__ __forceinline__ void row(int m, const int (&n)[NJ], float (&v)[NJ], int N) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (n[j] < N)
                C[(size_t)m * ldc + n[j]] = tanhf(__fadd_rn(v[j] + bias[n[j]], __fmul_rn(t, Sk[(size_t)m * lds + n[j]])));
    }
};

// Backward through an activation: C[m,n] = acc * act'(H[m,n]); column sums -> bias gradient.
//   MASK_RELU: act' = (H > 0);  MASK_TANH: act' = 1 - H^2;  MASK_NONE: 1.  ACCUM: C += ...
enum Mask { MASK_NONE = 0, MASK_RELU = 1, MASK_TANH = 2 };
template <int MASK, bool COLSUM, bool ACCUM>
struct EpiMaskStore {
    static constexpr bool kColSum = COLSUM;
    float* C; int ldc; const float* H; int ldh; float* colsum; float mul;
    __device__ float* colsum_dst() { return colsum; }
    template <int NJ>
    __device__ __forceinline__ void row(int m, const int (&n)[NJ], float (&v)[NJ], int N) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (n[j] < N) {
                float x = v[j];
                if (MASK == MASK_RELU) x = (H[(size_t)m * ldh + n[j]] > 0.f) ? x : 0.f;
                if (MASK == MASK_TANH) { const float h = H[(size_t)m * ldh + n[j]]; x = x * (1.f - h * h); }
                if (ACCUM) x += C[(size_t)m * ldc + n[j]];
                C[(size_t)m * ldc + n[j]] = x;
                v[j] = x;
            } else {
                v[j] = 0.f;
            }
        }
    }
};

// torch.optim.Adam (single-tensor CPU form) on one element, optional Polyak blend into a target.
struct AdamScalars {
    float lr_over_bc1;     // step_size = lr / (1 - beta1^t)
    float bc2_sqrt;        // sqrt(1 - beta2^t)
};
__device__ __forceinline__ float adam_element(float w, float g, float& m, float& v, const AdamScalars& s) {
    m = fmaf(0.1f, __fsub_rn(g, m), m);                                   // lerp_(g, 1 - beta1): ATen's vector path is an fma
    v = __fadd_rn(__fmul_rn(v, 0.999f), __fmul_rn(__fmul_rn(0.001f, g), g));   // mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), s.bc2_sqrt), 1e-8f);     // (sqrt(v) / sqrt(bc2)).add_(eps)
    return __fadd_rn(w, __fdiv_rn(__fmul_rn(-s.lr_over_bc1, m), denom));        // addcdiv_: w + (value * m) / denom
}

// dW tile -> Adam on W (row-major [rows x ld], element (m,n) or transposed (n,m)), moments alongside,
// optional target blend t = t*(1-tau) + tau*w (SAC.update_target_q / DDPG.update_target_nets).
template <bool TRANSPOSED, bool POLYAK>
struct EpiAdam {
    static constexpr bool kColSum = false;
    float* W; float* Mo; float* Vo; float* T; int ld; AdamScalars s; float tau, one_minus_tau;
    __device__ float* colsum_dst() { return nullptr; }
    template <int NJ>
    __device__ __forceinline__ void row(int m, const int (&n)[NJ], float (&g)[NJ], int N) {
        float w[NJ], mo[NJ], vo[NJ], t[NJ];
        size_t idx[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            idx[j] = TRANSPOSED ? (size_t)n[j] * ld + m : (size_t)m * ld + n[j];
            if (n[j] < N) {
                w[j] = W[idx[j]]; mo[j] = Mo[idx[j]]; vo[j] = Vo[idx[j]];
                if (POLYAK) t[j] = T[idx[j]];
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (n[j] < N) {
                const float wn = adam_element(w[j], g[j], mo[j], vo[j], s);
                W[idx[j]] = wn; Mo[idx[j]] = mo[j]; Vo[idx[j]] = vo[j];
                if (POLYAK) T[idx[j]] = __fadd_rn(__fmul_rn(t[j], one_minus_tau), __fmul_rn(tau, wn));
            }
        }
    }
};

}  // namespace spp
