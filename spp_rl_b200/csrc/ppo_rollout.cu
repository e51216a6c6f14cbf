// SPP-PPO on-policy rollout, device-resident and vectorised over E environments: the body of A2C.collect_batch
// (rltoolkit/algorithms/a2c/a2c.py:144-184) with AcMOnPolicyTrainer.process_action (rltoolkit/acm/on_policy.py:34-53):
//   x = Memory.normalize(obs) -> Actor.act (mean = tanh(fc3 tanh(fc2 tanh(fc1 x))) * lim, action = mean + N(0,1) exp(log_scale),
//   log-prob; rltoolkit/basic_model.py:32-51) -> denormalise -> ACM(cat[x, target]) (quirk 18: the ACM sees the NORMALISED
//   observation) -> env.step -> store (obs, next_obs, action, logp, reward, done, end) + the ACM action.
// The reference steps ONE environment per Python iteration; here every CTA owns a slice of <= kRolloutRows environments and walks
// all T steps of them without leaving the SM: activations live in shared memory, the layers' transposed weights (47 KB at Walker2d
// shapes) are copied to shared memory once per launch (wider nets: through L1), and the [T][E] store is written step-major -- exactly
// the layout the update kernels read (traj_stride = E).  MuJoCo is unavailable offline: the environment is the synthetic one of the
// off-policy rollout (obs' = 0.98 obs + 0.1 tanh(a-mix) + 0.02 N(0,1); reward = obs'[0]; done with probability done_prob; time-limit
// truncation at max_ep_len as in a2c.py:168-171; reset to 0.1 N(0,1)).  Noise comes from Philox, or from injected tensors (tests).
// Dot products accumulate k ascending with one fmaf per product inside a k-slice (one slice = the FFMA tiles' order of ppo_act_kernel).
#include "ppo_rollout.h"

#include "common.cuh"
#include "update_kernel.cuh"      // NORM_* indices

namespace spp {

constexpr int kRolloutRows = 32;

// out[r][n] = act(sum_k in[r][k] Wt[k][n] + b[n] (+ t * skip[r][n])) (* scale[n]).  A work item is (column n, group of 8 rows); when the CTA
// has more threads than items the contraction is split over KS adjacent lanes (each a k-slice, summed by a shuffle tree), so that
// the serial k-chain per thread -- the latency of a rollout step -- shrinks with the number of environments a CTA owns.
template <int ACT, bool SCALE, bool SKIP>
__device__ __forceinline__ void dense_rows(const float* __restrict__ in, int ldin, int K, const float* __restrict__ Wt, int ldt,
                                           const float* __restrict__ bias, int N, float* __restrict__ out, int ldout, int R,
                                           const float* __restrict__ scale = nullptr, float* __restrict__ out_pre = nullptr,
                                           const float* __restrict__ skip = nullptr, int ldskip = 0, float t = 0.f) {
    const int items = ((R + 7) / 8) * N;
    int KS = 1;
    while (KS < 8 && items * KS * 2 <= (int)blockDim.x && K / (KS * 2) >= 4) KS *= 2;
    const int Kc = ((K / 4 + KS - 1) / KS) * 4;      // k per slice, a multiple of 4 (K is: padded layouts, pad weights are zero)
    for (int base = 0; base < items * KS; base += blockDim.x) {
        const int w = base + threadIdx.x;
        const bool active = w < items * KS;
        const int idx = active ? w / KS : 0, ks = w % KS;
        const int n = idx % N, r0 = (idx / N) * 8;
        const int k_lo = ks * Kc, k_hi = active ? min(K, k_lo + Kc) : 0;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 2
        for (int k = k_lo; k < k_hi; k += 4) {
            float wv[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) wv[kk] = Wt[(size_t)(k + kk) * ldt + n];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(in + (size_t)(r0 + i) * ldin + k);      // rows >= R hold zeros
                acc[i] = fmaf(v.x, wv[0], acc[i]); acc[i] = fmaf(v.y, wv[1], acc[i]);
                acc[i] = fmaf(v.z, wv[2], acc[i]); acc[i] = fmaf(v.w, wv[3], acc[i]);
            }
        }
        for (int o = KS >> 1; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        if (!active || ks != 0) continue;
        const float b = bias[n];
        const float sc = SCALE ? scale[n] : 1.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (r0 + i >= R) break;
            float x = acc[i] + b;
            if (SKIP) x = __fadd_rn(x, __fmul_rn(t, skip[(size_t)(r0 + i) * ldskip + n]));
            if (ACT == 2) x = tanhf(x);
            if (out_pre) out_pre[(size_t)(r0 + i) * ldout + n] = x;
            out[(size_t)(r0 + i) * ldout + n] = SCALE ? x * sc : x;
        }
    }
}

// The same contraction with its work-item mapping computed ONCE per launch (the step loop used to redo the integer divisions of the
// mapping in every layer of every step: ~150 of the ~600 instructions a layer cost) and RG rows per work item instead of always 8 --
// a CTA that owns two environments carries two accumulators, not six zero rows.  Valid when one pass of the CTA covers the layer.
struct DenseMap { int n, r0, k_lo, k_hi, KS; bool lead, single; };
template <int RG>
__device__ __forceinline__ DenseMap dense_map(int K, int N, int R) {
    DenseMap m;
    const int items = ((R + RG - 1) / RG) * N;
    int KS = 1;
    while (KS < 8 && items * KS * 2 <= (int)blockDim.x && K / (KS * 2) >= 4) KS *= 2;
    const int Kc = ((K / 4 + KS - 1) / KS) * 4;
    m.single = items * KS <= (int)blockDim.x;
    const int w = threadIdx.x;
    const bool active = w < items * KS;
    const int idx = active ? w / KS : 0, ks = w % KS;
    m.n = idx % N; m.r0 = (idx / N) * RG; m.KS = KS;
    m.k_lo = ks * Kc; m.k_hi = active ? min(K, m.k_lo + Kc) : 0;
    m.lead = active && ks == 0;
    return m;
}
template <int ACT, bool SCALE, bool SKIP, int RG>
__device__ __forceinline__ void dense_mapped(const DenseMap& m, const float* __restrict__ in, int ldin, int K, const float* __restrict__ Wt, int ldt,
                                             const float* __restrict__ bias, int N, float* __restrict__ out, int ldout, int R,
                                             const float* __restrict__ scale = nullptr, float* __restrict__ out_pre = nullptr,
                                             const float* __restrict__ skip = nullptr, int ldskip = 0, float t = 0.f) {
    if (!m.single) { dense_rows<ACT, SCALE, SKIP>(in, ldin, K, Wt, ldt, bias, N, out, ldout, R, scale, out_pre, skip, ldskip, t); return; }
    float acc[RG];
#pragma unroll
    for (int i = 0; i < RG; ++i) acc[i] = 0.f;
    const float* wp = Wt + m.n;
    const float* ip = in + (size_t)m.r0 * ldin;
#pragma unroll 2
    for (int k = m.k_lo; k < m.k_hi; k += 4) {
        float wv[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) wv[kk] = wp[(size_t)(k + kk) * ldt];      // shared memory (staged weights) or global
#pragma unroll
        for (int i = 0; i < RG; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(ip + (size_t)i * ldin + k);      // rows >= R hold zeros
            acc[i] = fmaf(v.x, wv[0], acc[i]); acc[i] = fmaf(v.y, wv[1], acc[i]);
            acc[i] = fmaf(v.z, wv[2], acc[i]); acc[i] = fmaf(v.w, wv[3], acc[i]);
        }
    }
    for (int o = m.KS >> 1; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < RG; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (!m.lead) return;
    const float b = bias[m.n];
    const float sc = SCALE ? scale[m.n] : 1.f;
#pragma unroll
    for (int i = 0; i < RG; ++i) {
        if (m.r0 + i >= R) break;
        float x = acc[i] + b;
        if (SKIP) x = __fadd_rn(x, __fmul_rn(t, skip[(size_t)(m.r0 + i) * ldskip + m.n]));
        if (ACT == 2) x = tanhf(x);
        if (out_pre) out_pre[(size_t)(m.r0 + i) * ldout + m.n] = x;
        out[(size_t)(m.r0 + i) * ldout + m.n] = SCALE ? x * sc : x;
    }
}

template <int RG>
__device__ __forceinline__ void ppo_rollout_body(const PpoRolloutArgs& a) {
    extern __shared__ __align__(16) float sm[];
    const int ob = a.L.ob, ldo = a.L.ldo, ac = a.ac, lda = a.lda, ldm = 2 * ldo;
    const int e0 = blockIdx.x * a.rows_per_cta;
    const int R = min(a.rows_per_cta, a.E - e0);
    if (R <= 0) return;
    // shared-memory activations, every row buffer padded to 8-row groups (zero rows beyond R)
    float* xin = sm;                               // [32][ldm]   [x | target]
    float* h1 = xin + kRolloutRows * ldm;          // [32][64]
    float* h2 = h1 + kRolloutRows * kPpoHidden;    // [32][64]
    float* mean = h2 + kRolloutRows * kPpoHidden;  // [32][ldo]
    float* m1 = mean + kRolloutRows * ldo;         // [32][ldm1]
    float* m2 = m1 + kRolloutRows * a.ldm1;        // [32][ldm2]
    float* ms = m2 + kRolloutRows * a.ldm2;        // [32][ldm2]  BasicAcM skip
    float* pa = ms + kRolloutRows * a.ldm2;        // [32][lda]
    float* obs = pa + kRolloutRows * lda;          // [32][ldo]   raw observation of every environment
    int* eplen = reinterpret_cast<int*>(obs + kRolloutRows * ldo);      // [32]
    int* flags = eplen + kRolloutRows;                                  // [32]
    float* mixv = reinterpret_cast<float*>(flags + kRolloutRows);       // [32]
    const int total = kRolloutRows * (ldm + 2 * kPpoHidden + ldo + a.ldm1 + 2 * a.ldm2 + lda + ldo);
    for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = 0.f;
    // the transposed weights and biases of every layer of a step, staged ONCE in shared memory (47 KB at Walker2d shapes): through
    // L1 they hit 83 % of the time (the step's stores stream through the same cache), and every miss is an L2 round trip inside the
    // serial k-chain of a layer
    float* wcur = mixv + kRolloutRows;
    auto stage_layer = [&](const float* net, const LayerDesc& l, const float*& wt, const float*& b) {
        if (!a.stage_weights) { wt = net + l.off_wt; b = net + l.off_b; return; }      // too large for shared memory: through L1
        const int nw = l.ld * l.ld_t;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) wcur[i] = net[l.off_wt + i];
        wt = wcur; wcur += (nw + 3) & ~3;
        for (int i = threadIdx.x; i < l.rows; i += blockDim.x) wcur[i] = net[l.off_b + i];
        b = wcur; wcur += (l.rows + 3) & ~3;
    };
    const float *aW1, *aB1, *aW2, *aB2, *aW3, *aB3, *mW1, *mB1, *mW2, *mB2, *mW3, *mB3, *mWs = nullptr, *mBs = nullptr;
    {
        const LayerDesc& l0 = a.L.actor.L[0]; const LayerDesc& l1 = a.L.actor.L[1]; const LayerDesc& l2 = a.L.actor.L[2];
        stage_layer(a.actor, l0, aW1, aB1); stage_layer(a.actor, l1, aW2, aB2); stage_layer(a.actor, l2, aW3, aB3);
        stage_layer(a.acm, a.acm_desc.L[0], mW1, mB1); stage_layer(a.acm, a.acm_desc.L[1], mW2, mB2); stage_layer(a.acm, a.acm_desc.L[2], mW3, mB3);
        if (a.acm_kind != ACM_MLP) stage_layer(a.acm, a.acm_desc.L[3], mWs, mBs);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * ob; i += blockDim.x) obs[(i / ob) * ldo + i % ob] = a.state[(size_t)(e0 + i / ob) * ldo + i % ob];
    for (int i = threadIdx.x; i < R; i += blockDim.x) eplen[i] = a.ep_len[e0 + i];
    __syncthreads();

    const LayerDesc& l0 = a.L.actor.L[0]; const LayerDesc& l1 = a.L.actor.L[1]; const LayerDesc& l2 = a.L.actor.L[2]; const LayerDesc& l3 = a.L.actor.L[3];
    const float* nsub = a.norm + NORM_NSUB * ldo; const float* ndiv = a.norm + NORM_NDIV * ldo;
    const float* doff = a.norm + NORM_DOFF * ldo; const float* dsc = a.norm + NORM_DSCALE * ldo; const float* lim = a.norm + NORM_LIM * ldo;
    const float* ls = a.actor + l3.off_w;
    const float kLogSqrt2Pi = 0.918938533204672741780329736406f;
    const NetDesc& M = a.acm_desc;
    // work-item mappings of the six (seven with BasicAcM's skip) dense layers of a step
    const DenseMap ma1 = dense_map<RG>(ldo, kPpoHidden, R), ma2 = dense_map<RG>(kPpoHidden, kPpoHidden, R), ma3 = dense_map<RG>(kPpoHidden, ob, R);
    const DenseMap mm1 = dense_map<RG>(ldm, a.hm1, R), mm2 = dense_map<RG>(a.ldm1, a.hm2, R), mm3 = dense_map<RG>(a.ldm2, ac, R);
    const DenseMap mms = dense_map<RG>(ldm, a.hm2, R);

    for (int t = 0; t < a.T; ++t) {
        const size_t row0 = (size_t)t * a.E + e0;
        // 1. normalise (Memory.normalize, rltoolkit/buffer/memory.py:170-176) and store obs
        for (int i = threadIdx.x; i < R * ob; i += blockDim.x) {
            const int r = i / ob, j = i % ob;
            const float o = obs[r * ldo + j];
            float v = __fdiv_rn(__fsub_rn(o, nsub[j]), ndiv[j]);
            if (a.clamp) v = fminf(fmaxf(v, -10.f), 10.f);
            xin[r * ldm + j] = v;
            a.x[(row0 + r) * ldo + j] = v;
            if (a.raw_obs) a.raw_obs[(row0 + r) * ldo + j] = o;
        }
        __syncthreads();
        // 2. actor
        dense_mapped<2, false, false, RG>(ma1, xin, ldm, ldo, aW1, l0.ld_t, aB1, kPpoHidden, h1, kPpoHidden, R);
        __syncthreads();
        dense_mapped<2, false, false, RG>(ma2, h1, kPpoHidden, kPpoHidden, aW2, l1.ld_t, aB2, kPpoHidden, h2, kPpoHidden, R);
        __syncthreads();
        dense_mapped<2, true, false, RG>(ma3, h2, kPpoHidden, kPpoHidden, aW3, l2.ld_t, aB3, ob, mean, ldo, R, lim);
        __syncthreads();
        // 3. sample, denormalised target: one thread per (environment, column); the log-prob terms are left in `mean` (in place)
        //    and summed j ascending -- torch's sum(-1) order -- by the environment's thread in stage 5a
        for (int i = threadIdx.x; i < R * ob; i += blockDim.x) {
            const int r = i / ob, j = i % ob;
            float nz;
            if (a.noise_act) nz = a.noise_act[(row0 + r) * ob + j];
            else {
                const uint4 w = Philox::gen(a.seed, (uint64_t)t, (uint64_t)(e0 + r) * 128 + j);
                nz = normal_from_bits(w.x, w.y);
            }
            const float sd = expf(ls[j]);
            const float mu = mean[r * ldo + j];
            const float act = __fadd_rn(mu, __fmul_rn(nz, sd));
            const float d = __fsub_rn(act, mu);
            mean[r * ldo + j] = __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.f, __fmul_rn(sd, sd))), logf(sd)), kLogSqrt2Pi);
            a.act[(row0 + r) * ldo + j] = act;
            xin[r * ldm + ldo + j] = a.denorm_out ? __fadd_rn(doff[j], __fmul_rn(act, dsc[j])) : act;
        }
        __syncthreads();
        // 4. ACM (AcM: tanh-tanh-tanh * action limit; BasicAcM: skip connection and learnable gains)
        if (a.acm_kind == ACM_MLP) {
            dense_mapped<2, false, false, RG>(mm1, xin, ldm, ldm, mW1, M.L[0].ld_t, mB1, a.hm1, m1, a.ldm1, R);
            __syncthreads();
            dense_mapped<2, false, false, RG>(mm2, m1, a.ldm1, a.ldm1, mW2, M.L[1].ld_t, mB2, a.hm2, m2, a.ldm2, R);
            __syncthreads();
            dense_mapped<2, true, false, RG>(mm3, m2, a.ldm2, a.ldm2, mW3, M.L[2].ld_t, mB3, ac, pa, lda, R, a.acm_lim);
        } else {
            dense_mapped<2, false, false, RG>(mm1, xin, ldm, ldm, mW1, M.L[0].ld_t, mB1, a.hm1, m1, a.ldm1, R);
            dense_mapped<0, false, false, RG>(mms, xin, ldm, ldm, mWs, M.L[3].ld_t, mBs, a.hm2, ms, a.ldm2, R);
            __syncthreads();
            dense_mapped<2, false, true, RG>(mm2, m1, a.ldm1, a.ldm1, mW2, M.L[1].ld_t, mB2, a.hm2, m2, a.ldm2, R, nullptr,
                                       nullptr, ms, a.ldm2, __ldg(a.acm + M.L[4].off_w));
            __syncthreads();
            dense_mapped<2, true, false, RG>(mm3, m2, a.ldm2, a.ldm2, mW3, M.L[2].ld_t, mB3, ac, pa, lda, R,
                                       a.acm + M.L[4].off_w + 4);
        }
        __syncthreads();
        // 5a. per environment: log-prob sum, action mix, episode bookkeeping (flags[r]: bit 0 done, bit 1 end)
        for (int r = threadIdx.x; r < R; r += blockDim.x) {
            const int e = e0 + r;
            float lp = 0.f;
            for (int j = 0; j < ob; ++j) lp += mean[r * ldo + j];
            a.logp[row0 + r] = lp;
            float mix = 0.f;
            for (int j = 0; j < ac; ++j) {
                const float v = pa[r * lda + j];
                if (a.aacm) a.aacm[(row0 + r) * lda + j] = v;
                mix += v * (0.3f + 0.1f * (float)j);
            }
            mixv[r] = tanhf(mix);
            float u_done;
            if (a.u_done) u_done = a.u_done[row0 + r];
            else { const uint4 wd = Philox::gen(a.seed ^ 0xD0Eull, (uint64_t)t, (uint64_t)e); u_done = (float)(wd.x >> 8) * (1.0f / 16777216.0f); }
            const int len = eplen[r] + 1;
            const bool done = u_done < a.done_prob;
            const bool end = done || len == a.max_ep_len;
            flags[r] = end ? 1 : 0;
            a.done[row0 + r] = (done && len != a.max_ep_len) ? 1.f : 0.f;      // a2c.py:170: done = False if ep_len == max_ep_len else done
            a.end[row0 + r] = (end || t == a.T - 1) ? 1.f : 0.f;               // the batch cuts every trajectory at its last step
            eplen[r] = end ? 0 : len;
        }
        __syncthreads();
        // 5b. environment step + store, one thread per (environment, column)
        for (int i = threadIdx.x; i < R * ob; i += blockDim.x) {
            const int r = i / ob, j = i % ob, e = e0 + r;
            float nz;
            if (a.noise_env) nz = a.noise_env[(row0 + r) * ob + j];
            else {
                const uint4 w = Philox::gen(a.seed ^ 0xE9Full, (uint64_t)t, (uint64_t)e * 128 + j);
                nz = normal_from_bits(w.x, w.y);
            }
            const float o = obs[r * ldo + j];
            const float nx = 0.98f * o + 0.1f * mixv[r] * (1.f - 0.01f * (float)j) + 0.02f * nz;
            if (j == 0) a.rew[row0 + r] = nx;
            float v = __fdiv_rn(__fsub_rn(nx, nsub[j]), ndiv[j]);
            if (a.clamp) v = fminf(fmaxf(v, -10.f), 10.f);
            a.xn[(row0 + r) * ldo + j] = v;
            if (a.raw_next) a.raw_next[(row0 + r) * ldo + j] = nx;
            float nxt = nx;
            if (flags[r]) {      // env.reset(): the next rollout of this environment starts here
                float rz;
                if (a.noise_reset) rz = a.noise_reset[(row0 + r) * ob + j];
                else {
                    const uint4 w = Philox::gen(a.seed ^ 0x5E7ull, (uint64_t)t, (uint64_t)e * 128 + j);
                    rz = normal_from_bits(w.x, w.y);
                }
                nxt = 0.1f * rz;
            }
            obs[r * ldo + j] = nxt;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < R * ob; i += blockDim.x) a.state[(size_t)(e0 + i / ob) * ldo + i % ob] = obs[(i / ob) * ldo + i % ob];
    for (int i = threadIdx.x; i < R; i += blockDim.x) a.ep_len[e0 + i] = eplen[i];
}

__global__ void __launch_bounds__(256, 2) ppo_rollout_kernel(const __grid_constant__ PpoRolloutArgs a) {
    if (a.rows_per_cta <= 2) ppo_rollout_body<2>(a);
    else if (a.rows_per_cta <= 4) ppo_rollout_body<4>(a);
    else ppo_rollout_body<8>(a);
}

size_t ppo_rollout_smem_bytes(const PpoRolloutArgs& a) {
    const int ldo = a.L.ldo;
    size_t w = 0;      // staged weights: W^T [ld x ld_t] + b [rows] of every layer, each padded to 4 floats
    auto add = [&](const LayerDesc& l) { w += ((size_t)l.ld * l.ld_t + 3) / 4 * 4 + ((size_t)l.rows + 3) / 4 * 4; };
    add(a.L.actor.L[0]); add(a.L.actor.L[1]); add(a.L.actor.L[2]);
    add(a.acm_desc.L[0]); add(a.acm_desc.L[1]); add(a.acm_desc.L[2]);
    if (a.acm_kind != ACM_MLP) add(a.acm_desc.L[3]);
    const size_t act = (size_t)kRolloutRows * (2 * ldo + 2 * kPpoHidden + ldo + a.ldm1 + 2 * a.ldm2 + a.lda + ldo) * 4 + 3 * kRolloutRows * 4 + 16;
    return a.stage_weights ? act + w * 4 : act;
}

cudaError_t launch_ppo_rollout(const PpoRolloutArgs& a_in, cudaStream_t s) {
    PpoRolloutArgs a = a_in;
    a.stage_weights = 1;
    if (ppo_rollout_smem_bytes(a) > 110 * 1024) a.stage_weights = 0;      // keep two CTAs per SM; wide observations (Ant) read the weights through L1
    const size_t smem = ppo_rollout_smem_bytes(a);
    cudaError_t e = cudaFuncSetAttribute(ppo_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (a.E + a.rows_per_cta - 1) / a.rows_per_cta;
    ppo_rollout_kernel<<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace spp
