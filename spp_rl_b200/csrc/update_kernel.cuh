// Fused SPP-SAC / SPP-DDPG update burst: one persistent CTA per agent runs `G` consecutive
// update steps (replay gather -> Q target -> critic step(s) -> policy step -> Polyak -> temperature)
// without returning to the host.  Reference semantics: SAC_AcM.update
// (rltoolkit/acm/off_policy/sac_acm.py:89-162), DDPG_AcM.update
// (rltoolkit/acm/off_policy/ddpg_acm.py:147-201) driven by DDPG.make_update
// (rltoolkit/algorithms/ddpg/ddpg.py:231-237) with BufferAcMOffPolicy.sample_batch
// (rltoolkit/buffer/replay_buffer.py:385-398,233-261).
#pragma once
#include <cstddef>

#include "gemm_tile.cuh"
#include "gemm_umma.cuh"
#include "layout.h"

namespace spp {

struct RingPtrs {
    const float* obs;        // [P][S][ldo]
    const int32_t* oidx;     // [P][S]
    const int32_t* nidx;     // [P][S]
    const float* act;        // [P][S][ldo]   state-target actions (may be null when acm_critic)
    const float* rew;        // [P][S]
    const uint8_t* done;     // [P][S]
    const float* aacm;       // [P][S][lda]
    int64_t S;
};

// explicit minibatches (host-facing update(obs, next_obs, action, reward, done, acm_action))
struct BatchPtrs {
    const float* obs;        // [P][G][B][ob]  dense
    const float* nobs;
    const float* act;        // [P][G][B][ob]  (may be null when acm_critic)
    const float* rew;        // [P][G][B]
    const int8_t* done;      // [P][G][B]
    const float* aacm;       // [P][G][B][ac]
    const float* acm_x;      // [P][G][Bacm][2 ob]  ACM regression inputs cat[obs, next_obs]   (acm training only)
    const float* acm_y;      // [P][G][Bacm][ac]    ACM regression targets
};

struct Hyper {
    float gamma, tau, one_minus_tau, custom_loss, target_entropy;
    double actor_lr, critic_lr, alpha_lr, acm_lr;
    int norm_closs;          // custom loss in normalised space (MSE(z, normalize(next_obs)))
    int norm_clamp;          // mean-std normalize clamps to +-10 (utils.standardize_and_clip)
    int obs_norm;            // sample_batch normalises obs / next_obs (ReplayBuffer._sample_batch, replay_buffer.py:246-248); ring gather only
};

// per-agent normalisation vectors, each ldo floats: [P][NORM_COUNT][ldo]
enum { NORM_DOFF = 0, NORM_DSCALE = 1, NORM_NSUB = 2, NORM_NDIV = 3, NORM_LIM = 4, NORM_COUNT = 5 };

struct UpdateArgs {
    Layout L;
    Hyper h;
    float* params;           // [P][params_size]
    float* mom_m;            // [P][train_size]
    float* mom_v;            // [P][train_size]
    float* scratch;          // [P][s.size]
    int* steps;              // [P][4]   Adam step counters: actor, critic_1, critic_2, acm
    double* alpha_state;     // [P][4]   log_alpha, m, v, step
    const float* norm;       // [P][NORM_COUNT][ldo]
    const float* acm_lim;    // [lda]    env action limit (AcM) -- BasicAcM uses its own t1
    RingPtrs ring;
    BatchPtrs batch;
    const int64_t* idx;      // [P][G][B] ring indices, or null
    const int64_t* ring_len; // [P] current_len, used when idx == null (device sampler)
    const float* eps;        // [P][G][2][B][ob] injected N(0,1) draws (target pass, policy pass), or null
    uint64_t seed;           // Philox key when eps / idx are generated on device
    uint64_t seq;            // burst sequence number (Philox stream)
    float* losses;           // [P][G][8] or null
    int G;
    int population;
    int batch_row_stride;    // rows between consecutive steps in the staged batch / index arrays (0 = L.B)
    int acm_last_rows;       // ACM regression: rows of the final step when it is a partial batch (0 = full)
    int acm_eval;            // ACM regression: forward + loss only (calculate_validation_loss), no optimiser step
    unsigned long long* timing;   // [G][kStageMarks] %globaltimer stamps of agent 0's stage boundaries (spp_update_stage_profile), or null
    unsigned int* progress;  // [P] updates completed per agent in THIS launch (zeroed by the host), or null.  Set when the population
                             // exceeds the grid: the burst then runs as (step, agent) work items interleaved over the CTAs instead of whole
                             // agents per CTA, see update_burst_kernel
    int use_umma;            // GEMM path of the 256-wide products: 1 tcgen05 3-pass tf32 split (fp32-accurate, default), 0 FFMA tiles,
                             // 2 tcgen05 single tf32 pass (reduced-precision variant, stated tolerance 1e-2)
};

constexpr int kStageMarks = 24;
enum { LOSS_CRITIC_1 = 0, LOSS_CRITIC_2 = 1, LOSS_ACTOR = 2, LOSS_PI = 3, LOSS_DIST = 4, LOSS_ALPHA = 5, LOSS_ALPHA_VALUE = 6, LOSS_COUNT = 8 };

struct Smem {
    float gemm[kUmmaSmemBytes / 4];   // operand pipeline of the tile GEMMs (tcgen05 path: four 32 KB swizzled planes; FFMA path: first 102 KB)
    float red[8 * 2 * kHidden];       // cross-warp column partials (8 warps x 512)
    float vecs[4 * kHidden];          // reduced vectors
    float zpad[(kUmmaBPlane - (8 * 2 * kHidden + 4 * kHidden) * 4) / 4];      // red | vecs | zpad = 32 KB: the raw-B landing zone of
                                      // umma_mainloop_z; red / vecs only carry data inside the row-wise stages, never across a wide product
    float small[64];
    AdamScalars adam[4];
    float alpha;                      // temperature as fp32 (Python float rounded when it meets fp32 tensors)
    uint32_t tmem_base;               // TMEM allocation of this CTA (tcgen05 path)
    UmmaCtx um;                       // tcgen05 pipeline state (shared, see gemm_umma.cuh)
    uint64_t mbar[kUmmaSlots];        // one mbarrier per pipeline slot: "the MMAs reading this half-plane have retired"
};
static_assert(kUmmaSmemBytes / 4 >= kGemmSmemFloats, "the FFMA pipeline aliases the tcgen05 stages");
constexpr size_t kSmemLaunchBytes = sizeof(Smem) + 1024;   // + slack to align the base to a swizzle atom (1024 B)
static_assert(offsetof(Smem, vecs) == offsetof(Smem, red) + sizeof(float) * 8 * 2 * kHidden && offsetof(Smem, zpad) == offsetof(Smem, vecs) + sizeof(float) * 4 * kHidden,
              "the landing zone is red | vecs | zpad, contiguous");
static_assert(offsetof(Smem, red) % 1024 == 0, "landing zone keeps the operand planes' bank alignment");
static_assert(kSmemLaunchBytes <= 232448, "227 KB of dynamic shared memory per CTA on sm_100");

__device__ __forceinline__ Smem& smem_struct(unsigned char* raw) {
    return *reinterpret_cast<Smem*>(raw + ((1024u - (umma::smem_u32(raw) & 1023u)) & 1023u));
}

// ------------------------------------------------------------------------------------------------
struct Ctx {
    const UpdateArgs& a;
    int agent;
    float* P;      // params of this agent
    float* Mm;     // moments
    float* Mv;
    float* S;      // scratch
    Smem& sm;
    UmmaCtx* um;   // tcgen05 pipeline state, or null -> FFMA tiles
    __device__ Ctx(const UpdateArgs& a_, int agent_, Smem& sm_, UmmaCtx* um_ = nullptr) : a(a_), agent(agent_), sm(sm_), um(um_) {
        P = a.params + (size_t)agent * a.L.params_size;
        Mm = a.mom_m + (size_t)agent * a.L.train_size;
        Mv = a.mom_v + (size_t)agent * a.L.train_size;
        S = a.scratch + (size_t)agent * a.L.s.size;
    }
    __device__ float* net(int id) const { return P + a.L.net_off[id]; }
    __device__ float* net_m(int id) const { return Mm + a.L.net_off[id]; }
    __device__ float* net_v(int id) const { return Mv + a.L.net_off[id]; }
    __device__ const float* normv(int which) const {
        return a.norm + ((size_t)agent * NORM_COUNT + which) * a.L.ldo;
    }
    __device__ float* vec(int which) const { return S + a.L.s.vec + which * a.L.Bp; }
    __device__ float* gvec(int which) const { return S + a.L.s.gvec + which * 512; }
};

// ---- tcgen05 state of a persistent CTA: TMEM accumulators (all 512 columns) and the slot barriers live for the whole launch
__device__ __forceinline__ UmmaCtx* umma_setup(Smem& sm, int path) {
    if (warp_id() == 0) umma::tmem_alloc<kUmmaTmemCols>(&sm.tmem_base);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kUmmaSlots; ++s) umma::mbar_init(sm.mbar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (threadIdx.x == 0) {
        sm.um.smem = reinterpret_cast<unsigned char*>(sm.gemm);
        sm.um.mbar = sm.mbar;
        sm.um.tmem = sm.tmem_base;
        sm.um.phase_bits = 0;
        sm.um.dbg = (path == 2) ? kUmmaSinglePass : 0u;
        sm.um.zraw = umma::smem_u32(sm.red);      // used by the kernels built with SPP_UMMA_LANDING_ZONE (the update bursts)
    }
    __syncthreads();
    return &sm.um;
}
__device__ __forceinline__ void umma_teardown(UmmaCtx* um) {
    umma::fence_before_sync();
    __syncthreads();
    if (warp_id() == 0) umma::tmem_dealloc<kUmmaTmemCols>(um->tmem);
}

// ---- Adam bookkeeping: bump the step of optimiser `opt` and publish its scalars (thread 0 only)
__device__ inline void adam_begin(const Ctx& c, int opt, double lr) {
    int* st = c.a.steps + (size_t)c.agent * 4 + opt;
    const int t = *st + 1;
    *st = t;
    const double bc1 = 1.0 - pow(0.9, (double)t);
    const double bc2 = 1.0 - pow(0.999, (double)t);
    c.sm.adam[opt].lr_over_bc1 = (float)(lr / bc1);
    c.sm.adam[opt].bc2_sqrt = (float)sqrt(bc2);
}

// Adam (+ optional Polyak) on a contiguous vector with gradient in shared or global memory.
__device__ inline void adam_vector(float* W, float* Mo, float* Vo, float* T, const float* grad, int n,
                                   const AdamScalars& s, float tau, float omt, bool grad_global) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const float g = grad_global ? __ldcg(grad + i) : grad[i];
        float m = Mo[i], v = Vo[i];
        const float wn = adam_element(W[i], g, m, v, s);
        W[i] = wn; Mo[i] = m; Vo[i] = v;
        if (T) T[i] = __fadd_rn(__fmul_rn(T[i], omt), __fmul_rn(tau, wn));
    }
}

// ---- [M x 256] GEMM C = A . B on tcgen05 (gemm_umma.cuh), or on the FFMA tiles when c.um is null.  Operand convention of
//      gemm_tile.cuh: A_KC = true: A is [M x K]; false (dW): A is [K x M]; B is always [K x N].
template <bool A_KC, class Epi>
__device__ __forceinline__ void gemm_big(const Ctx& c, const float* A, int lda, const float* B, int ldb, int M, int N, int K, Epi epi) {
    if (c.um && N == 256) gemm256_umma<A_KC, false, Epi>(A, lda, B, ldb, M, K, *c.um, epi);
    else gemm<BigTile, A_KC>(A, lda, B, ldb, M, N, K, c.sm.gemm, epi);
}

// ---- forward of one Linear through the tile GEMM: C = act(X W^T + b), reading the transposed weight copy
template <class Cfg, int ACT, bool SCALE>
__device__ inline void linear_fwd(const Ctx& c, const float* X, int ldx, int K, const float* net, const LayerDesc& l,
                                  float* C, int ldc, int B, const float* scale = nullptr, float* C2 = nullptr,
                                  int ldc2 = 0, float* relu_bits = nullptr, const float* dot_w = nullptr, float* dot_out = nullptr) {
    EpiBiasAct<ACT, SCALE, false> epi{C, ldc, net + l.off_b, scale, C2, ldc2, nullptr, 0, 0.f, reinterpret_cast<unsigned long long*>(relu_bits),
                                      dot_w, dot_out};
    if constexpr (Cfg::TN == 128) gemm_big<true>(c, X, ldx, net + l.off_wt, l.ld_t, B, l.rows, K, epi);
    else gemm<Cfg, true>(X, ldx, net + l.off_wt, l.ld_t, B, l.rows, K, c.sm.gemm, epi);
}

// ---- replay gather (warp per row): fills xo, xn, xc=[obs|action], r, notdone*gamma
__device__ inline void stage_gather(const Ctx& c, int g) {
    const Layout& L = c.a.L;
    const int B = L.B, ob = L.ob, ldo = L.ldo;
    float* xo = c.S + L.s.xo; float* xn = c.S + L.s.xn; float* xc = c.S + L.s.xc;
    float* vr = c.vec(VEC_R); float* vnd = c.vec(VEC_ND);
    const bool from_ring = (c.a.batch.obs == nullptr);
    // Two phases per block of 256 rows, so that the dependent loads (index -> obs_idx / next_obs_idx -> rows) of ALL rows are in
    // flight together instead of one row per warp at a time: (1) one thread per row resolves the three row pointers into shared
    // memory and writes reward / gamma (1 - done); (2) the rows are copied with the (row, column) pairs flattened over the CTA.
    const float** rp = reinterpret_cast<const float**>(c.sm.red);      // [3][kThreads] row pointers: obs, next_obs, action
    for (int r0 = 0; r0 < B; r0 += kThreads) {
        const int nr = min(kThreads, B - r0);
        const int r = r0 + threadIdx.x;
        if (threadIdx.x < nr) {
            const float *po, *pn, *pa; float rew; int done;
            if (from_ring) {
                const RingPtrs& R = c.a.ring;
                int64_t i;
                if (c.a.idx) {
                    i = c.a.idx[((size_t)c.agent * c.a.G + g) * B + r];
                } else {   // device sampler: uniform over [0, current_len)
                    const uint4 x = Philox::gen(c.a.seed ^ 0x9E3779B97F4A7C15ull, ((uint64_t)c.agent << 32) | (uint32_t)g,
                                                (c.a.seq << 20) | (uint32_t)r);
                    i = (int64_t)__umul64hi(((uint64_t)x.x << 32) | x.y, (uint64_t)c.a.ring_len[c.agent]);
                }
                const size_t base = (size_t)c.agent * R.S;
                po = R.obs + (base + R.oidx[base + i]) * ldo;
                pn = R.obs + (base + R.nidx[base + i]) * ldo;
                pa = L.acm_critic ? R.aacm + (base + i) * L.lda : R.act + (base + i) * ldo;
                rew = R.rew[base + i];
                done = R.done[base + i];
            } else {
                const BatchPtrs& Bt = c.a.batch;
                const size_t row = ((size_t)c.agent * c.a.G + g) * B + r;
                po = Bt.obs + row * ob; pn = Bt.nobs + row * ob;
                pa = L.acm_critic ? Bt.aacm + row * L.ac : Bt.act + row * ob;
                rew = Bt.rew[row];
                done = Bt.done[row];
            }
            rp[threadIdx.x] = po; rp[kThreads + threadIdx.x] = pn; rp[2 * kThreads + threadIdx.x] = pa;
            vr[r] = rew;
            vnd[r] = __fmul_rn(c.a.h.gamma, (float)(1 - done));   // gamma * (1 - done)
        }
        __syncthreads();
        const bool norm = from_ring && c.a.h.obs_norm;      // explicit minibatches arrive as sample_batch returned them
        const float* nsub = c.normv(NORM_NSUB); const float* ndiv = c.normv(NORM_NDIV);
        for (int e = threadIdx.x; e < nr * ob; e += kThreads) {
            const int rr = e / ob, j = e - rr * ob, row = r0 + rr;
            float o = rp[rr][j], n = rp[kThreads + rr][j];
            if (norm) {
                o = __fdiv_rn(__fsub_rn(o, nsub[j]), ndiv[j]); n = __fdiv_rn(__fsub_rn(n, nsub[j]), ndiv[j]);
                if (c.a.h.norm_clamp) { o = fminf(fmaxf(o, -10.f), 10.f); n = fminf(fmaxf(n, -10.f), 10.f); }
            }
            xo[row * ldo + j] = o;
            xn[row * ldo + j] = n;
            xc[row * L.ldc + j] = o;
        }
        for (int e = threadIdx.x; e < nr * L.act_dim; e += kThreads) {
            const int rr = e / L.act_dim, j = e - rr * L.act_dim;
            xc[(r0 + rr) * L.ldc + ldo + j] = rp[2 * kThreads + rr][j];
        }
        __syncthreads();
    }
}

// ---- SAC: sample from the actor heads (warp per row).  pass 0: target pass on next_obs, pass 1: policy pass on obs
__device__ inline void stage_sample(const Ctx& c, int g, int pass) {
    const Layout& L = c.a.L;
    const int B = L.B, ob = L.ob, ldo = L.ldo, lane = lane_id();
    const float* x = c.S + (pass == 0 ? L.s.xn : L.s.xo);
    const float* ml = c.S + L.s.ml;
    float* xm = c.S + L.s.xm; float* xcp = c.S + L.s.xcp;
    float* zt = c.S + L.s.zt; float* epsb = c.S + L.s.epsb;
    float* logp_out = c.vec(pass == 0 ? VEC_LOGPN : VEC_LOGP);
    const float* lim = c.normv(NORM_LIM); const float* doff = c.normv(NORM_DOFF); const float* dsc = c.normv(NORM_DSCALE);
    const float kLogSqrt2Pi = 0.918938533204672741780329736406f, kLog2 = 0.693147180559945309417232121458f;
    // A row is handled by a group of gs lanes (the smallest power of two >= ob, at least 4), so a warp works on 32 / gs rows at a
    // time (Hopper: 2, Pendulum: 8) instead of leaving the lanes >= ob idle.  The group-wide xor tree gives bit-identical sums to
    // a 32-lane tree over zero-padded lanes.
    int gs = 32;
    while (gs > 4 && (gs >> 1) >= ob) gs >>= 1;
    const int rpw = 32 / gs, grp = lane / gs, gl = lane % gs;
    for (int rb = warp_id() * rpw; rb < B; rb += kWarps * rpw) {
        const int r = rb + grp;
        const bool valid = r < B;
        float lp = 0.f, corr = 0.f;
        if (valid)
        for (int j = gl; j < ob; j += gs) {
            const float mu = ml[r * L.ldh + j];
            const float ls = fminf(fmaxf(ml[r * L.ldh + ob + j], -20.f), 2.f);
            const float sd = expf(ls);
            float e;
            if (c.a.eps) {
                e = c.a.eps[((((size_t)c.agent * c.a.G + g) * 2 + pass) * B + r) * ob + j];
            } else {
                const uint4 w = Philox::gen(c.a.seed, ((uint64_t)c.agent << 32) | ((uint32_t)g * 2 + pass),
                                            (c.a.seq << 24) | ((uint32_t)r * ob + j));
                e = normal_from_bits(w.x, w.y);
            }
            const float u = __fadd_rn(mu, __fmul_rn(e, sd));
            const float var = __fmul_rn(sd, sd);
            const float d = __fsub_rn(u, mu);
            lp += __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.f, var)), logf(sd)), kLogSqrt2Pi);
            corr += __fsub_rn(__fsub_rn(kLog2, u), softplus_t(-2.f * u));
            const float th = tanhf(u);
            const float z = __fmul_rn(th, lim[j]);
            const float zd = __fadd_rn(doff[j], __fmul_rn(z, dsc[j]));
            const float xv = x[r * ldo + j];
            xcp[r * L.ldc + j] = xv;
            if (L.acm_critic) { xm[r * L.ldm + j] = xv; xm[r * L.ldm + ldo + j] = zd; }
            else xcp[r * L.ldc + ldo + j] = zd;
            if (pass == 1) { zt[r * ldo + j] = th; epsb[r * ldo + j] = e; if (!L.acm_critic) xm[r * L.ldm + ldo + j] = zd; }
        }
        for (int o = gs >> 1; o > 0; o >>= 1) {
            lp += __shfl_xor_sync(0xffffffffu, lp, o);
            corr += __shfl_xor_sync(0xffffffffu, corr, o);
        }
        if (valid && gl == 0) logp_out[r] = __fsub_rn(lp, __fmul_rn(2.f, corr));
    }
}

// ---- DDPG: deterministic head post-processing (heads GEMM already applied tanh; zt holds tanh, ml holds tanh*lim)
__device__ inline void stage_ddpg_post(const Ctx& c, int pass) {
    const Layout& L = c.a.L;
    const int B = L.B, ob = L.ob, ldo = L.ldo, lane = lane_id();
    const float* x = c.S + (pass == 0 ? L.s.xn : L.s.xo);
    const float* ml = c.S + L.s.ml;
    float* xm = c.S + L.s.xm; float* xcp = c.S + L.s.xcp;
    const float* doff = c.normv(NORM_DOFF); const float* dsc = c.normv(NORM_DSCALE);
    for (int e = threadIdx.x; e < B * ob; e += kThreads) {      // (row, column) pairs flattened over the CTA
        const int r = e / ob, j = e - r * ob;
        const float zd = __fadd_rn(doff[j], __fmul_rn(ml[r * L.ldh + j], dsc[j]));
        const float xv = x[r * ldo + j];
        xcp[r * L.ldc + j] = xv;
        if (L.acm_critic) { xm[r * L.ldm + j] = xv; xm[r * L.ldm + ldo + j] = zd; }
        else { xcp[r * L.ldc + ldo + j] = zd; if (pass == 1) xm[r * L.ldm + ldo + j] = zd; }
    }
}

// ---- ACM forward (AcM: tanh-tanh-tanh*lim; BasicAcM: skip connection and learnable gains).  Writes the action
//      to `out` (the action block of xcp in the update step) and keeps tanh(fc3) in tm3 for the backward.
__device__ inline void acm_forward(const Ctx& c, float* out, int ldout) {
    const Layout& L = c.a.L;
    const float* acm = c.net(NET_ACM);
    float* S = c.S;
    const int B = L.B;
    if (L.acm_kind == ACM_MLP) {
        linear_fwd<NarrowTile, ACT_TANH, false>(c, S + L.s.xm, L.ldm, L.ldm, acm, L.acm.L[0], S + L.s.hm1, L.ldm1, B);
        __syncthreads();
        linear_fwd<NarrowTile, ACT_TANH, false>(c, S + L.s.hm1, L.ldm1, L.ldm1, acm, L.acm.L[1], S + L.s.hm2, L.ldm2, B);
        __syncthreads();
        linear_fwd<NarrowTile, ACT_TANH, true>(c, S + L.s.hm2, L.ldm2, L.ldm2, acm, L.acm.L[2], out, ldout, B,
                                               c.a.acm_lim, S + L.s.tm3, L.lda);
        __syncthreads();
    } else {
        // h = tanh(fc1 x); s = fc21 x
        linear_fwd<NarrowTile, ACT_TANH, false>(c, S + L.s.xm, L.ldm, L.ldm, acm, L.acm.L[0], S + L.s.hm1, L.ldm1, B);
        linear_fwd<NarrowTile, ACT_NONE, false>(c, S + L.s.xm, L.ldm, L.ldm, acm, L.acm.L[3], S + L.s.hms, L.ldm2, B);
        __syncthreads();
        // h1 = tanh(fc2 h + t * s)
        {
            const LayerDesc& l = L.acm.L[1];
            EpiBiasAct<ACT_TANH, false, true> epi{S + L.s.hm2, L.ldm2, acm + l.off_b, nullptr, nullptr, 0,
                                                  S + L.s.hms, L.ldm2, acm[L.acm.L[4].off_w]};
            gemm<NarrowTile, true>(S + L.s.hm1, L.ldm1, acm + l.off_wt, l.ld_t, B, l.rows, L.ldm1, c.sm.gemm, epi);
        }
        __syncthreads();
        linear_fwd<NarrowTile, ACT_TANH, true>(c, S + L.s.hm2, L.ldm2, L.ldm2, acm, L.acm.L[2], out, ldout, B,
                                               acm + L.acm.L[4].off_w + 4, S + L.s.tm3, L.lda);
        __syncthreads();
    }
}

// ---- ACM backward to its input (frozen ACM: no weight gradients).  In: dxc action block.  Out: dxm.
__device__ inline void acm_backward_dx(const Ctx& c) {
    const Layout& L = c.a.L;
    const float* acm = c.net(NET_ACM);
    float* S = c.S;
    const int B = L.B, ac = L.ac;
    const float* scale = (L.acm_kind == ACM_MLP) ? c.a.acm_lim : acm + L.acm.L[4].off_w + 4;
    // d3 = da * scale * (1 - t3^2)
    for (int e = threadIdx.x; e < B * ac; e += kThreads) {
        const int r = e / ac, j = e % ac;
        const float t3 = S[L.s.tm3 + r * L.lda + j];
        const float da = S[L.s.dxc + r * L.ldc + L.ldo + j];
        S[L.s.dm3 + r * L.lda + j] = __fmul_rn(__fmul_rn(da, scale[j]), __fsub_rn(1.f, __fmul_rn(t3, t3)));
    }
    __syncthreads();
    {   // d2 = (d3 W3) * (1 - h2^2)      [B x hm2], K = ac
        const LayerDesc& l = L.acm.L[2];
        EpiMaskStore<MASK_TANH, false, false> epi{S + L.s.dm2, L.ldm2, S + L.s.hm2, L.ldm2, nullptr};
        gemm<NarrowTile, true>(S + L.s.dm3, L.lda, acm + l.off_w, l.ld, B, L.hm2, ac, c.sm.gemm, epi);
    }
    __syncthreads();
    if (L.acm_kind == ACM_BASIC) {   // ds = d2 * t
        const float t = acm[L.acm.L[4].off_w];
        for (int e = threadIdx.x; e < B * L.hm2; e += kThreads) {
            const int r = e / L.hm2, j = e % L.hm2;
            S[L.s.dms + r * L.ldm2 + j] = __fmul_rn(S[L.s.dm2 + r * L.ldm2 + j], t);
        }
    }
    {   // d1 = (d2 W2) * (1 - h1^2)      [B x hm1], K = hm2
        const LayerDesc& l = L.acm.L[1];
        EpiMaskStore<MASK_TANH, false, false> epi{S + L.s.dm1, L.ldm1, S + L.s.hm1, L.ldm1, nullptr};
        gemm<NarrowTile, true>(S + L.s.dm2, L.ldm2, acm + l.off_w, l.ld, B, L.hm1, L.hm2, c.sm.gemm, epi);
    }
    __syncthreads();
    {   // dx = d1 W1 (+ ds W21)          [B x ldm], K = hm1
        const LayerDesc& l = L.acm.L[0];
        EpiMaskStore<MASK_NONE, false, false> epi{S + L.s.dxm, L.ldm, nullptr, 0, nullptr};
        gemm<NarrowTile, true>(S + L.s.dm1, L.ldm1, acm + l.off_w, l.ld, B, L.ldm, L.hm1, c.sm.gemm, epi);
        if (L.acm_kind == ACM_BASIC) {
            const LayerDesc& l2 = L.acm.L[3];
            EpiMaskStore<MASK_NONE, false, true> epi2{S + L.s.dxm, L.ldm, nullptr, 0, nullptr};
            gemm<NarrowTile, true>(S + L.s.dms, L.ldm2, acm + l2.off_w, l2.ld, B, L.ldm, L.hm2, c.sm.gemm, epi2);
        }
    }
    __syncthreads();
}

constexpr int kRB = 4;      // rows a warp handles at a time in the critic-head stages

// ---- critic hidden layers for `ncrit` critics: hc1 = relu(fc1 x), hc2 = relu(fc2 hc1).
//      fused_head: hc2 is never stored; the fc2 epilogue leaves its relu mask as bits (mk_hc2) and the 32 partial dot products
//      per row of q = hc2 . w3 (qpart) -- all the target pass and the policy pass need (stage_qtarget, stage_policy_head_bwd).
__device__ inline void critics_hidden(const Ctx& c, const float* X, const int* nets, int ncrit, bool fused_head) {
    const Layout& L = c.a.L;
    float* S = c.S;
    for (int i = 0; i < ncrit; ++i)
        linear_fwd<BigTile, ACT_RELU, false>(c, X, L.ldc, L.ldc, c.net(nets[i]), L.critic.L[0], S + L.s.hc1[i], kHidden, L.B, nullptr,
                                             nullptr, 0, S + L.s.mk_hc1[i]);
    __syncthreads();
    for (int i = 0; i < ncrit; ++i) {
        const float* net = c.net(nets[i]);
        if (fused_head)
            linear_fwd<BigTile, ACT_RELU, false>(c, S + L.s.hc1[i], kHidden, kHidden, net, L.critic.L[1], nullptr, kHidden, L.B, nullptr, nullptr, 0,
                                                 S + L.s.mk_hc2[i], net + L.critic.L[2].off_w, S + L.s.qpart[i]);
        else
            linear_fwd<BigTile, ACT_RELU, false>(c, S + L.s.hc1[i], kHidden, kHidden, net, L.critic.L[1], S + L.s.hc2[i], kHidden, L.B);
    }
    __syncthreads();
}

// q of `ncrit` critics for 4 rows r0..r0+3 from the partials the fc2 epilogue left: lane s holds partial s of a row, the warp tree adds
// them in a fixed order.  Result valid in every lane.
__device__ __forceinline__ void q_from_partials(const Ctx& c, int r0, int ncrit, const float (&b3)[2], float (&q)[2][kRB]) {
    const Layout& L = c.a.L;
    float p[2][kRB];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr)
            p[i][rr] = (i < ncrit && r0 + rr < L.B) ? c.S[L.s.qpart[i] + (size_t)(r0 + rr) * 32 + lane_id()] : 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr) q[i][rr] = warp_sum(p[i][rr]) + b3[i];
}

// ---- critic heads.  fc3 is a 256-vector, so q = hc2 . w3 + b3 is done row-wise: each warp takes kRB rows at a
//      time, issues all of their loads first (kRB x 8 coalesced 128-byte reads per critic), then reduces.

struct HeadRows {
    float h[2][kRB][8];     // hc2 values of this lane: critic i, row rr, column lane + 32 k
    float q[2][kRB];
};

__device__ __forceinline__ void head_rows_load(const Ctx& c, int r0, int ncrit, const float (&w3)[2][8], const float (&b3)[2],
                                               HeadRows& hr) {
    const Layout& L = c.a.L;
    const int lane = lane_id();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (i >= ncrit) break;
        const float* h2 = c.S + L.s.hc2[i];
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr) {
            const int r = r0 + rr;
#pragma unroll
            for (int k = 0; k < 8; ++k) hr.h[i][rr][k] = (r < L.B) ? h2[(size_t)r * kHidden + lane + 32 * k] : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (i >= ncrit) break;
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(hr.h[i][rr][k], w3[i][k], s);
            hr.q[i][rr] = warp_sum(s) + b3[i];
        }
    }
}

__device__ __forceinline__ void head_weights_load(const Ctx& c, const int* nets, int ncrit, float (&w3)[2][8], float (&b3)[2]) {
    const Layout& L = c.a.L;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        b3[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) w3[i][k] = 0.f;
        if (i >= ncrit) continue;
        const float* net = c.net(nets[i]);
        b3[i] = net[L.critic.L[2].off_b];
#pragma unroll
        for (int k = 0; k < 8; ++k) w3[i][k] = net[L.critic.L[2].off_w + lane_id() + 32 * k];
    }
}

// ---- Q target: y = r + gamma*(1-done) * (min_i q_i^targ - alpha*logp')      (SAC) / (q^targ) (DDPG)
__device__ inline void stage_qtarget(const Ctx& c, const int* tnets, int ncrit) {
    const Layout& L = c.a.L;
    float* y = c.vec(VEC_Y);
    const float* vr = c.vec(VEC_R); const float* vnd = c.vec(VEC_ND); const float* lpn = c.vec(VEC_LOGPN);
    float b3[2] = {0.f, 0.f};
    for (int i = 0; i < ncrit; ++i) b3[i] = c.net(tnets[i])[L.critic.L[2].off_b];
    for (int r0 = warp_id() * kRB; r0 < L.B; r0 += kWarps * kRB) {
        float qq[2][kRB];
        q_from_partials(c, r0, ncrit, b3, qq);
        if (lane_id() < kRB && r0 + lane_id() < L.B) {
            const int r = r0 + lane_id();
            float q = 0.f;
#pragma unroll
            for (int rr = 0; rr < kRB; ++rr)
                if (rr == lane_id()) q = (ncrit == 2) ? fminf(qq[0][rr], qq[1][rr]) : qq[0][rr];
            float inner = q;
            if (L.algo == ALGO_SAC) inner = __fsub_rn(q, __fmul_rn(c.sm.alpha, lpn[r]));
            y[r] = __fadd_rn(vr[r], __fmul_rn(vnd[r], inner));
        }
    }
}

// ---- policy pass through the critic heads (fused form): loss = mean(alpha logp - min_i q_i) (SAC) / mean(-q) (DDPG), dq = -1/B routed
//      to argmin_i q_i, and dz2_i[r, k] = (hc2_i[r, k] > 0) dq_i w3_i[k] rebuilt from the relu bits -- hc2 itself was never stored.
//      Lane l owns the 8 columns of thread tx = l % 16 of column half nh = l / 16 in the epilogue's mapping.
__device__ inline void stage_policy_head_bwd(const Ctx& c, const int* nets, int ncrit, float* loss_out) {
    const Layout& L = c.a.L;
    float* S = c.S;
    const int B = L.B, lane = lane_id(), warp = warp_id();
    const float* lp = c.vec(VEC_LOGP);
    const float gq = -1.0f / (float)B, invB = 1.0f / (float)B;
    const int nh = lane >> 4, tx = lane & 15, c0 = nh * 128 + 4 * tx;
    float4 w3a[2], w3b[2]; float b3[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        w3a[i] = w3b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i >= ncrit) continue;
        const float* net = c.net(nets[i]);
        b3[i] = net[L.critic.L[2].off_b];
        w3a[i] = ld4(net + L.critic.L[2].off_w + c0);
        w3b[i] = ld4(net + L.critic.L[2].off_w + c0 + 64);
    }
    float lsum = 0.f;
    for (int r0 = warp * kRB; r0 < B; r0 += kWarps * kRB) {
        float qq[2][kRB];
        q_from_partials(c, r0, ncrit, b3, qq);
        unsigned long long wd[2][kRB];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int rr = 0; rr < kRB; ++rr) {
                const int r = r0 + rr;
                wd[i][rr] = 0ull;
                if (i < ncrit && r < B)
                    wd[i][rr] = reinterpret_cast<const unsigned long long*>(S + L.s.mk_hc2[i])[(size_t)((r >> 7) * 2 + nh) * kThreads + (r & 15) * 16 + tx];
            }
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr) {
            const int r = r0 + rr;
            if (r >= B) break;
            float dq[2] = {0.f, 0.f};
            if (ncrit == 2) {
                const float q0 = qq[0][rr], q1 = qq[1][rr];
                const float tie = (q0 == q1) ? 0.5f * gq : 0.f;
                dq[0] = (q0 < q1 ? gq : 0.f) + tie;
                dq[1] = (q1 < q0 ? gq : 0.f) + tie;
                const float qm = fminf(q0, q1);
                lsum += (lane == 0) ? (L.algo == ALGO_SAC ? __fsub_rn(__fmul_rn(c.sm.alpha, lp[r]), qm) : -qm) : 0.f;
            } else {
                dq[0] = gq;
                lsum += (lane == 0) ? -qq[0][rr] : 0.f;
            }
            const int sh = 8 * ((r & 127) >> 4);      // byte i = (r % 128) / 16 of the thread's word
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i >= ncrit) break;
                const unsigned int m8 = (unsigned int)(wd[i][rr] >> sh) & 255u;
                float* dz2 = S + L.s.dz2[i] + (size_t)r * kHidden;
                const float d = dq[i];
                st4(dz2 + c0, make_float4((m8 & 1u) ? __fmul_rn(d, w3a[i].x) : 0.f, (m8 & 2u) ? __fmul_rn(d, w3a[i].y) : 0.f,
                                          (m8 & 4u) ? __fmul_rn(d, w3a[i].z) : 0.f, (m8 & 8u) ? __fmul_rn(d, w3a[i].w) : 0.f));
                st4(dz2 + c0 + 64, make_float4((m8 & 16u) ? __fmul_rn(d, w3b[i].x) : 0.f, (m8 & 32u) ? __fmul_rn(d, w3b[i].y) : 0.f,
                                               (m8 & 64u) ? __fmul_rn(d, w3b[i].z) : 0.f, (m8 & 128u) ? __fmul_rn(d, w3b[i].w) : 0.f));
            }
        }
    }
    const float t = block_sum(lsum, c.sm.small);
    if (threadIdx.x == 0 && loss_out) loss_out[0] = t * invB;
    __syncthreads();
}

// ---- critic loss + backward through fc3 (+ Adam on fc3 / b3 / b2 of each critic).
//      MODE 0: critic update (dq = (2/B)(q - y); accumulates dW3, db3, db2 and applies Adam + Polyak to them)
//      MODE 1: policy pass   (dq = -1/B routed to argmin_i q_i; no weight gradients)
template <int MODE>
__device__ inline void stage_critic_head_bwd(const Ctx& c, const int* nets, const int* tnets, int ncrit, float* loss_out) {
    const Layout& L = c.a.L;
    float* S = c.S;
    const int B = L.B, lane = lane_id(), warp = warp_id();
    const float* y = c.vec(VEC_Y); const float* lp = c.vec(VEC_LOGP);
    const float invB = 1.0f / (float)B, norm2 = (float)(2.0 / (double)B);
    float acc_w3[2][8], acc_b2[2][8], acc_b3[2] = {0.f, 0.f}, lsum[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc_w3[i][k] = 0.f; acc_b2[i][k] = 0.f; }
    float w3[2][8], b3[2];
    head_weights_load(c, nets, ncrit, w3, b3);
    HeadRows hr;
    for (int r0 = warp * kRB; r0 < B; r0 += kWarps * kRB) {
        head_rows_load(c, r0, ncrit, w3, b3, hr);
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr) {
            const int r = r0 + rr;
            if (r >= B) break;
            float dq[2] = {0.f, 0.f};
            if (MODE == 0) {
                const float yr = y[r];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i >= ncrit) break;
                    const float diff = __fsub_rn(hr.q[i][rr], yr);
                    dq[i] = __fmul_rn(norm2, diff);
                    lsum[i] += (lane == 0) ? diff * diff : 0.f;
                }
            } else {
                const float gq = -invB;
                if (ncrit == 2) {
                    const float q0 = hr.q[0][rr], q1 = hr.q[1][rr];
                    const float tie = (q0 == q1) ? 0.5f * gq : 0.f;
                    dq[0] = (q0 < q1 ? gq : 0.f) + tie;
                    dq[1] = (q1 < q0 ? gq : 0.f) + tie;
                    const float qm = fminf(q0, q1);
                    lsum[0] += (lane == 0) ? (L.algo == ALGO_SAC ? __fsub_rn(__fmul_rn(c.sm.alpha, lp[r]), qm) : -qm) : 0.f;
                } else {
                    dq[0] = gq;
                    lsum[0] += (lane == 0) ? -hr.q[0][rr] : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i >= ncrit) break;
                float* dz2 = S + L.s.dz2[i] + (size_t)r * kHidden;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float h = hr.h[i][rr][k];
                    const float dz = (h > 0.f) ? __fmul_rn(dq[i], w3[i][k]) : 0.f;
                    dz2[lane + 32 * k] = dz;
                    if (MODE == 0) { acc_w3[i][k] = fmaf(dq[i], h, acc_w3[i][k]); acc_b2[i][k] += dz; }
                }
                if (MODE == 0) acc_b3[i] += dq[i];
            }
        }
    }
    // losses
    for (int i = 0; i < (MODE == 0 ? ncrit : 1); ++i) {
        const float t = block_sum(lsum[i], c.sm.small);
        if (threadIdx.x == 0 && loss_out) loss_out[i] = t * invB;
        __syncthreads();
    }
    if (MODE == 0) {
        // cross-warp reduction of the column partials, then Adam on fc3.weight, fc3.bias, fc2.bias
        for (int i = 0; i < ncrit; ++i) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                c.sm.red[warp * 512 + lane + 32 * k] = acc_w3[i][k];
                c.sm.red[warp * 512 + 256 + lane + 32 * k] = acc_b2[i][k];
            }
            if (lane == 0) c.sm.small[32 + warp] = acc_b3[i];
            __syncthreads();
            for (int n = threadIdx.x; n < 512; n += kThreads) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) s += c.sm.red[w * 512 + n];
                c.sm.vecs[n] = s;
            }
            if (threadIdx.x == 0) {
                float s = 0.f;
                for (int w = 0; w < kWarps; ++w) s += c.sm.small[32 + w];
                c.sm.vecs[512] = s;
            }
            __syncthreads();
            const int id = nets[i];
            float* net = c.net(id); float* nm = c.net_m(id); float* nv = c.net_v(id); float* tn = c.net(tnets[i]);
            const AdamScalars& as = c.sm.adam[1 + i];
            const LayerDesc& l3 = L.critic.L[2]; const LayerDesc& l2 = L.critic.L[1];
            adam_vector(net + l3.off_w, nm + l3.off_w, nv + l3.off_w, tn + l3.off_w, c.sm.vecs, kHidden, as, c.a.h.tau, c.a.h.one_minus_tau, false);
            adam_vector(net + l3.off_b, nm + l3.off_b, nv + l3.off_b, tn + l3.off_b, c.sm.vecs + 512, 1, as, c.a.h.tau, c.a.h.one_minus_tau, false);
            adam_vector(net + l2.off_b, nm + l2.off_b, nv + l2.off_b, tn + l2.off_b, c.sm.vecs + 256, kHidden, as, c.a.h.tau, c.a.h.one_minus_tau, false);
        }
    }
}

// ---- actor hidden layers + heads.  X: [B x ldo]
template <int ALGO>
__device__ inline void actor_forward(const Ctx& c, const float* X, int actor_net) {
    const Layout& L = c.a.L;
    float* S = c.S;
    const float* net = c.net(actor_net);
    linear_fwd<BigTile, ACT_RELU, false>(c, X, L.ldo, L.ldo, net, L.actor.L[0], S + L.s.ha1, kHidden, L.B, nullptr, nullptr, 0, S + L.s.mk_ha1);
    __syncthreads();
    linear_fwd<BigTile, ACT_RELU, false>(c, S + L.s.ha1, kHidden, kHidden, net, L.actor.L[1], S + L.s.ha2, kHidden, L.B, nullptr, nullptr, 0,
                                         S + L.s.mk_ha2);
    __syncthreads();
    if (ALGO == ALGO_SAC)
        linear_fwd<NarrowTile, ACT_NONE, false>(c, S + L.s.ha2, kHidden, kHidden, net, L.actor.L[2], S + L.s.ml, L.ldh, L.B);
    else   // tanh(fc3) * lim ; tanh kept in zt
        linear_fwd<NarrowTile, ACT_TANH, true>(c, S + L.s.ha2, kHidden, kHidden, net, L.actor.L[2], S + L.s.ml, L.ldh, L.B,
                                               c.normv(NORM_LIM), S + L.s.zt, L.ldo);
    __syncthreads();
}

struct RolloutArgs {
    UpdateArgs u;              // L.B = E (environments per agent); scratch sized for E rows
    // explicit form (host-facing noise_action + process_action):
    const float* in_obs;       // [P][E][ob]   raw observations
    const float* in_noise;     // [P][E][ob]   N(0,1) exploration noise (torch.randn in the reference), or null -> Philox
    const float* in_eps;       // [P][E][ob]   N(0,1) of the SAC actor's rsample, or null -> Philox
    float* out_target;         // [P][E][ob]   state target handed to the ACM (what the ring stores as "action")
    float* out_action;         // [P][E][ac]   ACM action handed to the environment
    // device-resident synthetic form:
    float* env_state;          // [P][E][ldo]  current observation of every environment, or null (explicit form)
    float* w_obs; int32_t* w_oidx; int32_t* w_nidx; float* w_act; float* w_rew; uint8_t* w_done; uint8_t* w_end; float* w_aacm;
    const int64_t* obs_cur;    // [P] obs-row cursor at launch (row of env 0's CURRENT observation)
    const int64_t* ts_cur;     // [P] timestep cursor at launch
    int steps;                 // consecutive environment steps per launch (device form)
    int random_phase;          // frames < random_frames: target = lim * N(0,1) (initial_act)
    int obs_norm;              // normalise the actor / ACM input (replay_buffer.normalize gate)
    int denormalize_out;       // denormalize_actor_out
    float act_noise;
};

}  // namespace spp
