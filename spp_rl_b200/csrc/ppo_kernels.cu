// SPP-PPO kernels (see ppo_kernels.cuh).  Reference semantics:
//   A2C.update_critic / calculate_q_val   rltoolkit/algorithms/a2c/a2c.py:186-225,247-265
//   PPO.calculate_gae                     rltoolkit/algorithms/ppo/ppo.py:117-150
//   AdvantageDataset (normalisation)      rltoolkit/algorithms/ppo/advantage_dataset.py:9-16
//   PPO_AcM.update_actor_acm, _clip_loss  rltoolkit/acm/on_policy.py:164-216, rltoolkit/algorithms/ppo/ppo.py:194-204
//   Actor / Critic nets                   rltoolkit/basic_model.py:7-77
#include "ppo_kernels.cuh"
#include "update_kernel.cuh"      // NORM_* slots

namespace spp {

__device__ __forceinline__ float group8_sum(float v) {   // sum over the 8 lanes that share a row in MidTile (tx = lane % 8)
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// Column sums of a MidTile/any-tile register block over the CTA, added (single writer) to dst[n0 + c].
template <class Cfg>
__device__ __forceinline__ void tile_colsum_add(const float (&cs)[Cfg::NJ], int n0, int N, float* dst, float* smem) {
    const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
#pragma unroll
    for (int j = 0; j < Cfg::NJ; ++j) smem[ty * Cfg::TN + col_of<Cfg>(j, tx)] = cs[j];
    __syncthreads();
    for (int c = threadIdx.x; c < Cfg::TN; c += kThreads) {
        float s = 0.f;
#pragma unroll 4
        for (int t = 0; t < Cfg::TY; ++t) s += smem[t * Cfg::TN + c];
        if (n0 + c < N) dst[n0 + c] += s;
    }
    __syncthreads();
}

// ---- epilogue of the critic's fc2 GEMM (one MidTile spans all 64 hidden units): h2 stays in registers.
//   forward-only (q == null): V[m] = h2 . w3 + b3
//   training: dv = -(q - V) / Ntot; dz2 = dv * w3 * (1 - h2^2) stored; d w3, d b3, d b2 and the squared error accumulated
struct EpiCriticHead {
    const float* b2; const float* w3; float b3;
    const float* q; float* v_out; float* dz2; float inv_n;
    float* g_w3; float* g_b3; float* g_b2; float* loss_acc;
    template <class Cfg, bool A_KC>
    __device__ __forceinline__ void apply(float (&acc)[Cfg::MI][Cfg::NJ], int m0, int n0, int M, int N, float* smem) {
        const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
        float bv[Cfg::NJ], wv[Cfg::NJ], cs_b2[Cfg::NJ], cs_w3[Cfg::NJ];
#pragma unroll
        for (int j = 0; j < Cfg::NJ; ++j) {
            const int n = n0 + col_of<Cfg>(j, tx);
            bv[j] = b2[n]; wv[j] = w3[n]; cs_b2[j] = 0.f; cs_w3[j] = 0.f;
        }
        float sse = 0.f, sdv = 0.f;
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i) {
            const int m = m0 + row_of<Cfg, A_KC>(i, ty);
            float h[Cfg::NJ], dot = 0.f;
#pragma unroll
            for (int j = 0; j < Cfg::NJ; ++j) { h[j] = tanhf(acc[i][j] + bv[j]); dot = fmaf(h[j], wv[j], dot); }
            const float v = group8_sum(dot) + b3;      // every lane of the row group holds the full dot product
            if (m >= M) continue;
            if (v_out && tx == 0) v_out[m] = v;
            if (q) {
                const float diff = __fsub_rn(q[m], v);
                const float dv = -__fmul_rn(diff, inv_n);
                if (tx == 0) { sse = fmaf(diff, diff, sse); sdv += dv; }
                float dz[Cfg::NJ];
#pragma unroll
                for (int j = 0; j < Cfg::NJ; ++j) {
                    dz[j] = __fmul_rn(__fmul_rn(dv, wv[j]), __fsub_rn(1.f, __fmul_rn(h[j], h[j])));
                    cs_b2[j] += dz[j];
                    cs_w3[j] = fmaf(dv, h[j], cs_w3[j]);
                }
#pragma unroll
                for (int g = 0; g < Cfg::NJ / 4; ++g)
                    st4(dz2 + (size_t)m * kPpoHidden + n0 + col_of<Cfg>(4 * g, tx), make_float4(dz[4 * g], dz[4 * g + 1], dz[4 * g + 2], dz[4 * g + 3]));
            }
        }
        if (q) {
            tile_colsum_add<Cfg>(cs_b2, n0, N, g_b2, smem);
            tile_colsum_add<Cfg>(cs_w3, n0, N, g_w3, smem);
            const float t1 = block_sum(sse, smem);
            __syncthreads();
            const float t2 = block_sum(sdv, smem);
            if (threadIdx.x == 0) { *loss_acc += t1; *g_b3 += t2; }
            __syncthreads();
        }
    }
};

// plain store of a dW tile into the CTA's partial-gradient arena (natural [rows x ld] layout)
struct EpiStorePartial {
    float* G; int ld; bool accumulate;      // accumulate: add to what an earlier row block of this CTA left there (fixed order)
    template <class Cfg, bool A_KC>
    __device__ __forceinline__ void apply(float (&acc)[Cfg::MI][Cfg::NJ], int m0, int n0, int M, int N, float*) {
        const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i) {
            const int m = m0 + row_of<Cfg, A_KC>(i, ty);
            if (m >= M) continue;
#pragma unroll
            for (int g = 0; g < Cfg::NJ / 4; ++g) {
                const int n = n0 + col_of<Cfg>(4 * g, tx);
                if (n >= N) continue;
                float4 v = make_float4(acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]);
                if (accumulate) {
                    const float4 o = ld4(G + (size_t)m * ld + n);
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                st4(G + (size_t)m * ld + n, v);
            }
        }
    }
};

// Rows a CTA pushes through the whole layer chain before it moves on: the [rows x 64] intermediates (h1, dz2, dz1) of one block
// are 128 KB each and are re-used block after block, so they live in L2 (148 CTAs x 384 KB) instead of streaming through HBM --
// at config-4 sizes (56 k rows per CTA) the chain used to write and re-read ~10 GB per optimiser step.
constexpr int kPpoRowBlock = 512;

constexpr int kPpoCtasPerSm = 2;      // 8 warps per SM leave the FMA pipe 42 % busy (ncu): two CTAs (2 x 102 KB, <= 128 registers) share an SM
struct Chunk { int64_t r0; int rows; };
__device__ __forceinline__ Chunk my_chunk(int64_t n, int rows_per_cta) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    int64_t rows = n - r0;
    if (rows > rows_per_cta) rows = rows_per_cta;
    if (rows < 0) rows = 0;
    return Chunk{r0, (int)rows};
}

// ------------------------------------------------------------------------------------------------ critic
// mode 0: V(x) -> v, V(xn) -> nv, q = r + gamma (1 - done) nv        (calculate_q_val + the state values GAE needs)
// mode 1: q only (the target pass at the top of each of A2C.update_critic's outer iterations)
__global__ void __launch_bounds__(kThreads, kPpoCtasPerSm) ppo_critic_values_kernel(const __grid_constant__ PpoArgs a) {
    extern __shared__ __align__(16) float smem[];
    const Chunk ck = my_chunk(a.d.N, a.rows_per_cta);
    if (ck.rows == 0) return;
    const LayerDesc& l0 = a.L.critic.L[0]; const LayerDesc& l1 = a.L.critic.L[1]; const LayerDesc& l2 = a.L.critic.L[2];
    float* h1 = a.s.h1 + ck.r0 * kPpoHidden;      // the first kPpoRowBlock rows of the CTA's own region, re-used by every block
    for (int pass = (a.mode == 1 ? 1 : 0); pass < 2; ++pass) {
        for (int b0 = 0; b0 < ck.rows; b0 += kPpoRowBlock) {
            const int rows = min(kPpoRowBlock, ck.rows - b0);
            const float* X = (pass == 0 ? a.d.x : a.d.xn) + (ck.r0 + b0) * a.L.ldo;
            EpiBiasAct<ACT_TANH, false, false> e1{h1, kPpoHidden, a.critic + l0.off_b, nullptr, nullptr, 0, nullptr, 0, 0.f};
            gemm<MidTile, true>(X, a.L.ldo, a.critic + l0.off_wt, l0.ld_t, rows, kPpoHidden, a.L.ldo, smem, e1);
            __syncthreads();
            EpiCriticHead e2{a.critic + l1.off_b, a.critic + l2.off_w, a.critic[l2.off_b], nullptr, (pass == 0 ? a.d.v : a.d.nv) + ck.r0 + b0,
                             nullptr, 0.f, nullptr, nullptr, nullptr, nullptr};
            gemm<MidTile, true>(h1, kPpoHidden, a.critic + l1.off_wt, l1.ld_t, rows, kPpoHidden, kPpoHidden, smem, e2);
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < ck.rows; i += kThreads) {
        const int64_t r = ck.r0 + i;
        a.d.q[r] = __fadd_rn(a.d.rew[r], __fmul_rn(__fmul_rn(a.h.gamma, __fsub_rn(1.f, a.d.done[r])), a.d.nv[r]));
    }
}

__global__ void __launch_bounds__(kThreads, kPpoCtasPerSm) ppo_critic_grad_kernel(const __grid_constant__ PpoArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* part = a.part + (size_t)blockIdx.x * a.part_stride;
    float* scal = a.scal + (size_t)blockIdx.x * PS_COUNT;
    for (int i = threadIdx.x; i < a.L.critic.size; i += kThreads) part[i] = 0.f;
    if (threadIdx.x < PS_COUNT) scal[threadIdx.x] = 0.f;
    __syncthreads();
    const Chunk ck = my_chunk(a.d.N, a.rows_per_cta);
    if (ck.rows == 0) return;
    const LayerDesc& l0 = a.L.critic.L[0]; const LayerDesc& l1 = a.L.critic.L[1]; const LayerDesc& l2 = a.L.critic.L[2];
    float* h1 = a.s.h1 + ck.r0 * kPpoHidden; float* dz2 = a.s.dz2 + ck.r0 * kPpoHidden; float* dz1 = a.s.dz1 + ck.r0 * kPpoHidden;
    for (int b0 = 0; b0 < ck.rows; b0 += kPpoRowBlock) {      // the scratch of the CTA's first block is re-used by every block
        const int rows = min(kPpoRowBlock, ck.rows - b0);
        const float* X = a.d.x + (ck.r0 + b0) * a.L.ldo;
        EpiBiasAct<ACT_TANH, false, false> e1{h1, kPpoHidden, a.critic + l0.off_b, nullptr, nullptr, 0, nullptr, 0, 0.f};
        gemm<MidTile, true>(X, a.L.ldo, a.critic + l0.off_wt, l0.ld_t, rows, kPpoHidden, a.L.ldo, smem, e1);
        __syncthreads();
        EpiCriticHead e2{a.critic + l1.off_b, a.critic + l2.off_w, a.critic[l2.off_b], a.d.q + ck.r0 + b0, nullptr, dz2, 1.0f / (float)a.d.Ntot,
                         part + l2.off_w, part + l2.off_b, part + l1.off_b, scal + PS_LOSS};
        gemm<MidTile, true>(h1, kPpoHidden, a.critic + l1.off_wt, l1.ld_t, rows, kPpoHidden, kPpoHidden, smem, e2);
        __syncthreads();
        EpiMaskStore<MASK_TANH, true, false> e3{dz1, kPpoHidden, h1, kPpoHidden, part + l0.off_b};
        gemm<MidTile, true>(dz2, kPpoHidden, a.critic + l1.off_w, l1.ld, rows, kPpoHidden, kPpoHidden, smem, e3);
        __syncthreads();
        EpiStorePartial g2{part + l1.off_w, l1.ld, b0 > 0};
        gemm<SmallTile, false>(dz2, kPpoHidden, h1, kPpoHidden, kPpoHidden, kPpoHidden, rows, smem, g2);
        EpiStorePartial g1{part + l0.off_w, l0.ld, b0 > 0};
        gemm<SmallTile, false>(dz1, kPpoHidden, X, a.L.ldo, kPpoHidden, a.L.ldo, rows, smem, g1);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ actor
// One minibatch (rows already gathered into a.b): clipped-ratio loss, entropy bonus, partial gradients, KL sum.
__global__ void __launch_bounds__(kThreads, kPpoCtasPerSm) ppo_actor_grad_kernel(const __grid_constant__ PpoArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* part = a.part + (size_t)blockIdx.x * a.part_stride;
    float* scal = a.scal + (size_t)blockIdx.x * PS_COUNT;
    for (int i = threadIdx.x; i < a.L.actor.size; i += kThreads) part[i] = 0.f;
    if (threadIdx.x < PS_COUNT) scal[threadIdx.x] = 0.f;
    __syncthreads();
    const Chunk ck = my_chunk(a.b.n, a.rows_per_cta);
    if (ck.rows == 0) return;
    const int ob = a.L.ob, ldo = a.L.ldo;
    const LayerDesc& l0 = a.L.actor.L[0]; const LayerDesc& l1 = a.L.actor.L[1]; const LayerDesc& l2 = a.L.actor.L[2]; const LayerDesc& l3 = a.L.actor.L[3];
    const float* X = a.b.x + ck.r0 * ldo;
    float* h1 = a.s.h1 + ck.r0 * kPpoHidden; float* h2 = a.s.h2 + ck.r0 * kPpoHidden;
    float* dz2 = a.s.dz2 + ck.r0 * kPpoHidden; float* dz1 = a.s.dz1 + ck.r0 * kPpoHidden;
    float* mean = a.s.mean + ck.r0 * ldo; float* t3 = a.s.t3 + ck.r0 * ldo; float* d3 = a.s.d3 + ck.r0 * ldo;
    const float* lim = a.norm + NORM_LIM * ldo;
    // forward
    EpiBiasAct<ACT_TANH, false, false> e1{h1, kPpoHidden, a.actor + l0.off_b, nullptr, nullptr, 0, nullptr, 0, 0.f};
    gemm<MidTile, true>(X, ldo, a.actor + l0.off_wt, l0.ld_t, ck.rows, kPpoHidden, ldo, smem, e1);
    __syncthreads();
    EpiBiasAct<ACT_TANH, false, false> e2{h2, kPpoHidden, a.actor + l1.off_b, nullptr, nullptr, 0, nullptr, 0, 0.f};
    gemm<MidTile, true>(h1, kPpoHidden, a.actor + l1.off_wt, l1.ld_t, ck.rows, kPpoHidden, kPpoHidden, smem, e2);
    __syncthreads();
    EpiBiasAct<ACT_TANH, true, false> e3{mean, ldo, a.actor + l2.off_b, lim, t3, ldo, nullptr, 0, 0.f};
    gemm<NarrowTile, true>(h2, kPpoHidden, a.actor + l2.off_wt, l2.ld_t, ck.rows, ob, kPpoHidden, smem, e3);
    __syncthreads();
    // head: log-prob under the new policy, ratio, clipped loss and its gradient w.r.t. the mean and log_scale
    {
        const float* ls = a.actor + l3.off_w;
        const float invB = 1.0f / (float)a.b.n_mean;
        const float kLogSqrt2Pi = 0.918938533204672741780329736406f;
        float g_ls[32];      // per-lane partial of d log_scale (ob <= 128 -> up to 4 columns per lane... kept simple: ob <= 32*? see below)
        // one thread per row; column accumulators are reduced through shared memory in a fixed order afterwards
        float s_loss = 0.f, s_kl = 0.f, s_dist = 0.f;
        float* col = smem;                       // [kThreads][?] would be too large: accumulate per column via two-phase below
        (void)col; (void)g_ls;
        for (int i = threadIdx.x; i < ck.rows; i += kThreads) {
            const int64_t r = ck.r0 + i;
            float lp = 0.f;
            for (int j = 0; j < ob; ++j) {
                const float sd = expf(ls[j]);
                const float d = __fsub_rn(a.b.act[r * ldo + j], mean[(size_t)i * ldo + j]);
                lp += __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.f, __fmul_rn(sd, sd))), logf(sd)), kLogSqrt2Pi);
            }
            a.s.newlogp[r] = lp;
            const float old = a.b.logp[r], adv = a.b.adv[r];
            const float ratio = expf(__fsub_rn(lp, old));
            const float lo = 1.f - a.h.epsilon, hi = 1.f + a.h.epsilon;
            const float clipped = fminf(fmaxf(ratio, lo), hi);
            const float s1 = __fmul_rn(ratio, adv), s2 = __fmul_rn(clipped, adv);
            s_loss += a.a2c ? -__fmul_rn(old, adv) : -fminf(s1, s2);      // A2C: mean(-logp * adv) with the recorded log-prob (a2c.py:279)
            s_kl += __fsub_rn(old, lp);
            const float g = -invB;
            const float tie = (s1 == s2) ? 0.5f * g : 0.f;
            const float g1 = (s1 < s2 ? g : 0.f) + tie, g2 = (s2 < s1 ? g : 0.f) + tie;
            const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
            const float dratio = __fadd_rn(__fmul_rn(g1, adv), __fmul_rn(__fmul_rn(g2, adv), in_range));
            float dlogp = __fmul_rn(dratio, ratio);
            if (a.a2c) dlogp = __fmul_rn(g, adv);      // A2C: the recorded log-prob's graph, taken at these very weights (no ratio)
            for (int j = 0; j < ob; ++j) {
                const float sd = expf(ls[j]);
                const float var = __fmul_rn(sd, sd);
                const float d = __fsub_rn(a.b.act[r * ldo + j], mean[(size_t)i * ldo + j]);
                const float dmean = __fmul_rn(dlogp, __fdiv_rn(d, var));
                const float t = t3[(size_t)i * ldo + j];
                d3[(size_t)i * ldo + j] = __fmul_rn(__fmul_rn(dmean, lim[j]), __fsub_rn(1.f, __fmul_rn(t, t)));
                // d logp / d log_scale = d^2 / var - 1 ; stash the per-row term in `mean` (no longer needed) for the column pass
                mean[(size_t)i * ldo + j] = __fmul_rn(dlogp, __fsub_rn(__fdiv_rn(__fmul_rn(d, d), var), 1.f));
                if (a.h.custom_loss != 0.f) {
                    float la = a.b.act[r * ldo + j], ln = a.b.xn[r * ldo + j];               // both already in the loss's space ...
                    if (a.mode == 1) {      // ... except in the A2C form, whose log-prob above needs the normalised action
                        const float* doff = a.norm + NORM_DOFF * ldo; const float* dsc = a.norm + NORM_DSCALE * ldo;
                        la = __fadd_rn(doff[j], __fmul_rn(la, dsc[j])); ln = __fadd_rn(doff[j], __fmul_rn(ln, dsc[j]));
                    }
                    const float dd = __fsub_rn(la, ln);
                    s_dist = fmaf(dd, dd, s_dist);
                }
            }
        }
        const float t_loss = block_sum(s_loss, smem);
        __syncthreads();
        const float t_kl = block_sum(s_kl, smem);
        __syncthreads();
        const float t_dist = block_sum(s_dist, smem);
        __syncthreads();
        if (threadIdx.x == 0) { scal[PS_LOSS] = t_loss; scal[PS_KL] = t_kl; scal[PS_DIST] = t_dist; }
        // column passes (deterministic): d log_scale[j] = sum_rows term[j]; d b3[j] = sum_rows d3[j]
        for (int j = threadIdx.x; j < ob; j += kThreads) {
            float sl = 0.f, sb = 0.f;
            for (int i = 0; i < ck.rows; ++i) { sl += mean[(size_t)i * ldo + j]; sb += d3[(size_t)i * ldo + j]; }
            part[l3.off_w + j] = sl;
            part[l2.off_b + j] = sb;
        }
        __syncthreads();
    }
    // backward
    EpiMaskStore<MASK_TANH, true, false> b2{dz2, kPpoHidden, h2, kPpoHidden, part + l1.off_b};
    gemm<MidTile, true>(d3, ldo, a.actor + l2.off_w, l2.ld, ck.rows, kPpoHidden, ob, smem, b2);
    __syncthreads();
    EpiMaskStore<MASK_TANH, true, false> b1{dz1, kPpoHidden, h1, kPpoHidden, part + l0.off_b};
    gemm<MidTile, true>(dz2, kPpoHidden, a.actor + l1.off_w, l1.ld, ck.rows, kPpoHidden, kPpoHidden, smem, b1);
    __syncthreads();
    EpiStorePartial g3{part + l2.off_w, l2.ld};
    gemm<SmallTile, false>(d3, ldo, h2, kPpoHidden, ob, kPpoHidden, ck.rows, smem, g3);
    EpiStorePartial g2{part + l1.off_w, l1.ld};
    gemm<SmallTile, false>(dz2, kPpoHidden, h1, kPpoHidden, kPpoHidden, kPpoHidden, ck.rows, smem, g2);
    EpiStorePartial g1{part + l0.off_w, l0.ld};
    gemm<SmallTile, false>(dz1, kPpoHidden, X, ldo, kPpoHidden, ldo, ck.rows, smem, g1);
}

// ------------------------------------------------------------------------------------------------ on-policy act
// Rollout step of A2C.collect_batch (rltoolkit/algorithms/a2c/a2c.py:165-166): x = Memory.normalize(obs) (already in
// a.b.x), mean = tanh(fc3(tanh(fc2(tanh(fc1 x))))) * lim, action = mean + noise * exp(log_scale) (Normal.sample),
// logp = Independent(Normal).log_prob(action) (rltoolkit/basic_model.py:32-51).  Outputs: a.b.act (sampled target,
// normalised space), a.b.logp, a.b.xn (denormalised target handed to the ACM, rltoolkit/acm/on_policy.py:46-47).
// `a.b.adv` holds nothing here; the noise comes in through a.s.d3.
__global__ void __launch_bounds__(kThreads, 1) ppo_act_kernel(const __grid_constant__ PpoArgs a) {
    extern __shared__ __align__(16) float smem[];
    const Chunk ck = my_chunk(a.b.n, a.rows_per_cta);
    if (ck.rows == 0) return;
    const int ob = a.L.ob, ldo = a.L.ldo;
    const LayerDesc& l0 = a.L.actor.L[0]; const LayerDesc& l1 = a.L.actor.L[1]; const LayerDesc& l2 = a.L.actor.L[2]; const LayerDesc& l3 = a.L.actor.L[3];
    const float* X = a.b.x + ck.r0 * ldo;
    float* h1 = a.s.h1 + ck.r0 * kPpoHidden; float* h2 = a.s.h2 + ck.r0 * kPpoHidden;
    float* mean = a.s.mean + ck.r0 * ldo;
    const float* lim = a.norm + NORM_LIM * ldo; const float* doff = a.norm + NORM_DOFF * ldo; const float* dsc = a.norm + NORM_DSCALE * ldo;
    EpiBiasAct<ACT_TANH, false, false> e1{h1, kPpoHidden, a.actor + l0.off_b, nullptr, nullptr, 0, nullptr, 0, 0.f};
    gemm<MidTile, true>(X, ldo, a.actor + l0.off_wt, l0.ld_t, ck.rows, kPpoHidden, ldo, smem, e1);
    __syncthreads();
    EpiBiasAct<ACT_TANH, false, false> e2{h2, kPpoHidden, a.actor + l1.off_b, nullptr, nullptr, 0, nullptr, 0, 0.f};
    gemm<MidTile, true>(h1, kPpoHidden, a.actor + l1.off_wt, l1.ld_t, ck.rows, kPpoHidden, kPpoHidden, smem, e2);
    __syncthreads();
    EpiBiasAct<ACT_TANH, true, false> e3{mean, ldo, a.actor + l2.off_b, lim, nullptr, 0, nullptr, 0, 0.f};
    gemm<NarrowTile, true>(h2, kPpoHidden, a.actor + l2.off_wt, l2.ld_t, ck.rows, ob, kPpoHidden, smem, e3);
    __syncthreads();
    const float* ls = a.actor + l3.off_w;
    const float kLogSqrt2Pi = 0.918938533204672741780329736406f;
    for (int i = threadIdx.x; i < ck.rows; i += kThreads) {
        const int64_t r = ck.r0 + i;
        float lp = 0.f;
        for (int j = 0; j < ob; ++j) {
            const float sd = expf(ls[j]);
            const float mu = mean[(size_t)i * ldo + j];
            const float act = __fadd_rn(mu, __fmul_rn(a.s.d3[r * ldo + j], sd));
            const float d = __fsub_rn(act, mu);
            lp += __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.f, __fmul_rn(sd, sd))), logf(sd)), kLogSqrt2Pi);
            a.b.act[r * ldo + j] = act;
            a.b.xn[r * ldo + j] = (a.mode == 1) ? __fadd_rn(doff[j], __fmul_rn(act, dsc[j])) : act;
        }
        a.b.logp[r] = lp;
    }
}

// ------------------------------------------------------------------------------------------------ reduce / Adam
__global__ void ppo_reduce_kernel(const float* __restrict__ part, int stride, int n_part, int n, float* __restrict__ out,
                                  const float* __restrict__ scal, float* __restrict__ gscal) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        out[i] = slot_sum(part, (size_t)stride, n_part, i);
    }
    if (blockIdx.x == 0 && threadIdx.x < PS_COUNT) {
        float s = 0.f;
        for (int p = 0; p < n_part; ++p) s += scal[(size_t)p * PS_COUNT + threadIdx.x];
        gscal[threadIdx.x] = s;
    }
}

// reduce + record + Adam in one launch (single-GPU form of the library's loops; the data-parallel form is ppo_reduce_p2p_kernel)
__global__ void __launch_bounds__(256) ppo_reduce_step_kernel(const float* __restrict__ part, int stride, int n_part, int n, float* __restrict__ out,
                                                              const float* __restrict__ scal, StepArgs t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= stride + PS_COUNT) return;
    float g = 0.f;
    if (i < n) g = slot_sum(part, (size_t)stride, n_part, i);
    else if (i >= stride) g = slot_sum(scal, (size_t)PS_COUNT, n_part, i - stride);
    out[i] = g;
    if (i >= stride) { if (t.slog) t.slog[i - stride] = g; }
    else apply_step_element(t, i, g);
}

// Adam over every tensor of a net from the reduced gradient (natural layout); refreshes the transposed copies.
__global__ void ppo_adam_kernel(NetDesc d, float* __restrict__ W, float* __restrict__ Mo, float* __restrict__ Vo,
                                const float* __restrict__ G, AdamScalars s, float extra_ls_grad, int ls_layer) {
    for (int li = 0; li < d.n_layers; ++li) {
        const LayerDesc& l = d.L[li];
        const int nw = l.rows * l.ld;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nw + l.rows; i += gridDim.x * blockDim.x) {
            const bool is_b = i >= nw;
            if (is_b && l.off_wt < 0 && li == ls_layer) continue;      // pseudo layer (log_scale) has no bias
            const int o = is_b ? l.off_b + (i - nw) : l.off_w + i;
            if (!is_b && (i % l.ld) >= l.cols) continue;                // pad columns stay zero
            float g = G[o];
            if (li == ls_layer && !is_b) g += extra_ls_grad;            // - entropy_coef * d entropy / d log_scale (= 1 per dim)
            float m = Mo[o], v = Vo[o];
            const float wn = adam_element(W[o], g, m, v, s);
            W[o] = wn; Mo[o] = m; Vo[o] = v;
            if (!is_b && l.off_wt >= 0) W[l.off_wt + (size_t)(i % l.ld) * l.ld_t + (i / l.ld)] = wn;
        }
    }
}

// ------------------------------------------------------------------------------------------------ GAE, advantage statistics
// One thread per trajectory, reverse scan (ppo.py:139-148): done -> carry 0; non-terminal end -> bootstrap with V(next).  The carry is a
// serial fp32 chain, the five loads of a step are not: they are issued eight steps ahead of their use (the scan used to expose one
// L2 / HBM latency per time step: 2.6 ms for 2048-step trajectories whatever the number of environments).
__global__ void ppo_gae_kernel(PpoData d, PpoHyper h) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.n_traj) return;
    const int64_t start = d.traj_start[t], len = d.traj_len[t], st = d.traj_stride;
    constexpr int U = 8;
    float carry = 0.f;
    for (int64_t k0 = len - 1; k0 >= 0; k0 -= U) {
        float q[U], v[U], nv[U], dn[U], en[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t k = k0 - u;
            const int64_t r = start + (k >= 0 ? k : 0) * st;
            q[u] = d.q[r]; v[u] = d.v[r]; nv[u] = d.nv[r]; dn[u] = d.done[r]; en[u] = d.end[r];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t k = k0 - u;
            if (k < 0) break;
            const float delta = __fsub_rn(q[u], v[u]);
            if (dn[u] != 0.f) carry = delta;                                                         // 0 * discount + delta
            else if (en[u] != 0.f) carry = __fadd_rn((float)((double)nv[u] * h.discount_d), delta);   // python-float product
            else carry = __fadd_rn(__fmul_rn(carry, h.discount), delta);
            d.adv[start + k * st] = carry;
        }
    }
}

__global__ void ppo_adv_stats_kernel(const float* __restrict__ adv, int64_t n, double* __restrict__ stats) {
    __shared__ double sh[2][kThreads / 32];
    double s = 0.0, s2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)adv[i];
        s += v; s2 += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < blockDim.x / 32; ++w) { a += sh[0][w]; b += sh[1][w]; }
        stats[2 * blockIdx.x] = a; stats[2 * blockIdx.x + 1] = b;       // per-block partials; summed on the host in order
    }
}

__global__ void ppo_adv_apply_kernel(float* __restrict__ adv, int64_t n, float mean, float denom) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        adv[i] = __fdiv_rn(__fsub_rn(adv[i], mean), denom);          // (A - mean) / (std + 1.2e-7)
}

// gather the rows of a minibatch; `denorm`: hand the custom loss denormalised actions / next obs (norm_closs False)
__global__ void ppo_gather_kernel(PpoData d, PpoBatch b, const int64_t* __restrict__ perm, int ob, int ldo,
                                  const float* __restrict__ norm, int denorm) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float* doff = norm + NORM_DOFF * ldo; const float* dsc = norm + NORM_DSCALE * ldo;
    for (int64_t i = w; i < b.n; i += nw) {
        const int64_t r = perm[i];
        for (int j = lane; j < ob; j += 32) {
            b.x[i * ldo + j] = d.x[r * ldo + j];
            float av = d.act[r * ldo + j], nv = d.xn[r * ldo + j];
            if (denorm) { av = __fadd_rn(doff[j], __fmul_rn(av, dsc[j])); nv = __fadd_rn(doff[j], __fmul_rn(nv, dsc[j])); }
            b.act[i * ldo + j] = av;
            b.xn[i * ldo + j] = nv;
        }
        if (lane == 0) { b.logp[i] = d.logp[r]; b.adv[i] = d.adv[r]; }
    }
}

__global__ void ppo_normalize_rows_kernel(const float* __restrict__ raw, float* __restrict__ out, int64_t n, int ob, int ldo,
                                          const float* __restrict__ norm, int clamp) {
    const float* nsub = norm + NORM_NSUB * ldo; const float* ndiv = norm + NORM_NDIV * ldo;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * ob; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / ob; const int j = (int)(e % ob);
        float v = __fdiv_rn(__fsub_rn(raw[r * ob + j], nsub[j]), ndiv[j]);
        if (clamp) v = fminf(fmaxf(v, -10.f), 10.f);
        out[r * ldo + j] = v;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
static constexpr size_t kPpoSmem = (size_t)kGemmSmemFloats * sizeof(float);

cudaError_t launch_ppo_critic_values(const PpoArgs& a, int grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ppo_critic_values_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPpoSmem);
    if (e != cudaSuccess) return e;
    ppo_critic_values_kernel<<<grid, kThreads, kPpoSmem, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_ppo_critic_grad(const PpoArgs& a, int grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ppo_critic_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPpoSmem);
    if (e != cudaSuccess) return e;
    ppo_critic_grad_kernel<<<grid, kThreads, kPpoSmem, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_ppo_actor_grad(const PpoArgs& a, int grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ppo_actor_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPpoSmem);
    if (e != cudaSuccess) return e;
    ppo_actor_grad_kernel<<<grid, kThreads, kPpoSmem, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_ppo_act(const PpoArgs& a, int grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ppo_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPpoSmem);
    if (e != cudaSuccess) return e;
    ppo_act_kernel<<<grid, kThreads, kPpoSmem, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_ppo_reduce(const PpoArgs& a, int n_part, int n_elems, cudaStream_t s) {
    ppo_reduce_kernel<<<(n_elems + 255) / 256, 256, 0, s>>>(a.part, a.part_stride, n_part, n_elems, a.gbuf, a.scal, a.gscal);
    return cudaGetLastError();
}
cudaError_t launch_ppo_reduce_step(const PpoArgs& a, int n_part, int n_elems, const StepArgs& t, cudaStream_t s) {
    ppo_reduce_step_kernel<<<(a.part_stride + PS_COUNT + 255) / 256, 256, 0, s>>>(a.part, a.part_stride, n_part, n_elems, a.gbuf, a.scal, t);
    return cudaGetLastError();
}
cudaError_t launch_ppo_adam(const PpoArgs& a, int which_net, int step, double lr, cudaStream_t s) {
    const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = 1.0 - pow(0.999, (double)step);
    AdamScalars sc{(float)(lr / bc1), (float)sqrt(bc2)};
    if (which_net == 0)
        ppo_adam_kernel<<<32, 256, 0, s>>>(a.L.actor, a.actor, a.actor_m, a.actor_v, a.gbuf, sc, -a.h.entropy_coef, 3);
    else
        ppo_adam_kernel<<<32, 256, 0, s>>>(a.L.critic, a.critic, a.critic_m, a.critic_v, a.gbuf, sc, 0.f, -1);
    return cudaGetLastError();
}
cudaError_t launch_ppo_gae(const PpoArgs& a, cudaStream_t s) {
    ppo_gae_kernel<<<(a.d.n_traj + 127) / 128, 128, 0, s>>>(a.d, a.h);
    return cudaGetLastError();
}
cudaError_t launch_ppo_adv_stats(const PpoArgs& a, double* d_stats, int grid, cudaStream_t s) {
    ppo_adv_stats_kernel<<<grid, kThreads, 0, s>>>(a.d.adv, a.d.N, d_stats);
    return cudaGetLastError();
}
cudaError_t launch_ppo_adv_apply(const PpoArgs& a, float mean, float denom, int grid, cudaStream_t s) {
    ppo_adv_apply_kernel<<<grid, kThreads, 0, s>>>(a.d.adv, a.d.N, mean, denom);
    return cudaGetLastError();
}
cudaError_t launch_ppo_gather(const PpoArgs& a, const int64_t* d_perm, int denorm, int grid, cudaStream_t s) {
    ppo_gather_kernel<<<grid, kThreads, 0, s>>>(a.d, a.b, d_perm, a.L.ob, a.L.ldo, a.norm, denorm);
    return cudaGetLastError();
}
cudaError_t launch_ppo_normalize_rows(const float* raw, float* out, int64_t n, int ob, int ldo, const float* norm, int clamp,
                                      int grid, cudaStream_t s) {
    ppo_normalize_rows_kernel<<<grid, kThreads, 0, s>>>(raw, out, n, ob, ldo, norm, clamp);
    return cudaGetLastError();
}

}  // namespace spp
