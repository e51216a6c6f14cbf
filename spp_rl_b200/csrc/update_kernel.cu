// Fused off-policy SPP update burst kernel (SAC_AcM / DDPG_AcM), sm_100a.  See update_kernel.cuh.
#ifndef SPP_NO_LANDING_ZONE
#define SPP_UMMA_LANDING_ZONE      // wide products through umma_mainloop_z (gemm_umma.cuh); -DSPP_NO_LANDING_ZONE builds the two-slot loop for A/B
#endif
#include "update_kernel.cuh"

namespace spp {

// loss slots in Smem::small
constexpr int kLossBase = 8;

// ---- backward through the actor heads (warp per row); also custom (distance) loss and the SAC temperature step.
template <int ALGO>
__device__ inline void stage_actor_head_bwd(const Ctx& c) {
    const Layout& L = c.a.L;
    const Hyper& h = c.a.h;
    float* S = c.S;
    const int B = L.B, ob = L.ob, ldo = L.ldo, lane = lane_id(), warp = warp_id();
    const float* ml = S + L.s.ml; const float* zt = S + L.s.zt; const float* epsb = S + L.s.epsb;
    const float* xm = S + L.s.xm; const float* xn = S + L.s.xn;
    const float* din = L.acm_critic ? S + L.s.dxm : S + L.s.dxc;
    const int ldin = L.acm_critic ? L.ldm : L.ldc;
    float* dml = S + L.s.dml;
    const float* lim = c.normv(NORM_LIM); const float* dsc = c.normv(NORM_DSCALE);
    const float* nsub = c.normv(NORM_NSUB); const float* ndiv = c.normv(NORM_NDIV);
    const float* lp = c.vec(VEC_LOGP);
    const float invB = 1.0f / (float)B;
    const float g = __fmul_rn(c.sm.alpha, invB);
    const float norm_d = (float)(2.0 / ((double)B * (double)ob));
    float cb[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float dist_sum = 0.f, term_sum = 0.f, aterm_sum = 0.f;
    // lane groups as in stage_sample: a warp works on 32 / gs rows at a time
    int gs = 32;
    while (gs > 4 && (gs >> 1) >= ob) gs >>= 1;
    const int rpw = 32 / gs, grp = lane / gs, gl = lane % gs;
    for (int rb = warp * rpw; rb < B; rb += kWarps * rpw) {
        const int r = rb + grp;
        if (r >= B) continue;
        int k = 0;
        for (int j = gl; j < ob; j += gs, ++k) {
            const float th = zt[r * ldo + j];
            const float zd = xm[r * L.ldm + ldo + j];
            float dz = 0.f, dzd = 0.f;
            if (h.custom_loss != 0.f) {
                float pred, target;
                if (h.norm_closs) {
                    target = __fdiv_rn(__fsub_rn(xn[r * ldo + j], nsub[j]), ndiv[j]);
                    if (h.norm_clamp) target = fminf(fmaxf(target, -10.f), 10.f);
                    pred = __fmul_rn(th, lim[j]);
                } else {
                    target = xn[r * ldo + j];
                    pred = zd;
                }
                const float diff = __fsub_rn(pred, target);
                dist_sum = fmaf(diff, diff, dist_sum);
                const float gd = __fmul_rn(__fmul_rn(norm_d, diff), h.custom_loss);
                if (h.norm_closs) dz = gd; else dzd = gd;
            }
            dzd = __fadd_rn(dzd, din[r * ldin + ldo + j]);
            dz = __fadd_rn(dz, __fmul_rn(dzd, dsc[j]));
            if (ALGO == ALGO_SAC) {
                const float mu = ml[r * L.ldh + j];
                const float lsr = ml[r * L.ldh + ob + j];
                const float ls = fminf(fmaxf(lsr, -20.f), 2.f);
                const float sd = expf(ls);
                const float e = epsb[r * ldo + j];
                const float u = __fadd_rn(mu, __fmul_rn(e, sd));
                const float var = __fmul_rn(sd, sd);
                const float d = __fsub_rn(u, mu);
                const float two_var = __fmul_rn(2.f, var);
                const float ds = __fdiv_rn(-g, two_var);
                const float dd = __fmul_rn(__fmul_rn(ds, 2.f), d);
                const float d2var = __fdiv_rn(__fmul_rn(g, __fmul_rn(d, d)), __fmul_rn(two_var, two_var));
                const float dvar = __fmul_rn(2.f, d2var);
                float dstd = __fadd_rn(__fmul_rn(__fmul_rn(dvar, 2.f), sd), __fdiv_rn(-g, sd));
                float du = __fmul_rn(__fmul_rn(dz, lim[j]), __fsub_rn(1.f, __fmul_rn(th, th)));
                const float sg = 1.f / (1.f + expf(2.f * u));      // sigmoid(-2u)
                du = __fadd_rn(du, __fmul_rn(__fmul_rn(-g, 2.f), __fadd_rn(-1.f, __fmul_rn(2.f, sg))));
                du = __fadd_rn(du, dd);
                const float dmu = __fadd_rn(-dd, du);
                dstd = __fadd_rn(dstd, __fmul_rn(du, e));
                const float dls = __fmul_rn(dstd, sd);
                const float dlsr = (lsr >= -20.f && lsr <= 2.f) ? dls : 0.f;
                dml[r * L.ldh + j] = dmu;
                dml[r * L.ldh + ob + j] = dlsr;
                if (k < 4) { cb[0][k] += dmu; cb[1][k] += dlsr; }
            } else {
                const float d3 = __fmul_rn(__fmul_rn(dz, lim[j]), __fsub_rn(1.f, __fmul_rn(th, th)));
                dml[r * L.ldh + j] = d3;
                if (k < 4) cb[0][k] += d3;
            }
        }
        if (ALGO == ALGO_SAC && gl == 0) {
            const float term = __fsub_rn(-lp[r], h.target_entropy);
            term_sum += __fmul_rn(invB, term);
            aterm_sum += term;
        }
    }
    // head-bias gradients: cross-warp reduce -> gvec(GV_MISC)[0..heads)
    __syncthreads();
    {   // one slot of `heads` floats per lane group: (8 warps x rpw groups) x heads <= 1776 floats
        const int slot = warp * rpw + grp;
        int k = 0;
        for (int j = gl; j < ob; j += gs, ++k) {
            c.sm.red[slot * L.heads + j] = cb[0][k];
            if (ALGO == ALGO_SAC) c.sm.red[slot * L.heads + ob + j] = cb[1][k];
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < L.heads; n += kThreads) {
        float s = 0.f;
        for (int w = 0; w < kWarps * rpw; ++w) s += c.sm.red[w * L.heads + n];      // fixed order: deterministic
        c.gvec(GV_MISC)[n] = s;
    }
    const float dist_tot = block_sum(dist_sum, c.sm.small);
    __syncthreads();
    const float term_tot = (ALGO == ALGO_SAC) ? block_sum(term_sum, c.sm.small) : 0.f;
    __syncthreads();
    const float aterm_tot = (ALGO == ALGO_SAC) ? block_sum(aterm_sum, c.sm.small) : 0.f;
    if (threadIdx.x == 0) {
        float* ls = c.sm.small + kLossBase;
        const float dist = dist_tot / ((float)B * (float)ob);
        ls[LOSS_DIST] = dist;
        ls[LOSS_ACTOR] = (h.custom_loss != 0.f) ? __fadd_rn(ls[LOSS_PI], __fmul_rn(h.custom_loss, dist)) : ls[LOSS_PI];
        if (ALGO == ALGO_SAC) {
            // temperature: SAC.compute_alpha_loss (rltoolkit/algorithms/sac/sac.py:201-216) + Adam on the fp64 log_alpha
            double* as = c.a.alpha_state + (size_t)c.agent * 4;
            const double la = as[0];
            const double ea = exp(la);
            ls[LOSS_ALPHA] = __fmul_rn((float)ea, aterm_tot) * invB;
            const double gla = (double)term_tot * ea;
            const int t = (int)as[3] + 1;
            double m = as[1], v = as[2];
            m = m + 0.09999999999999998 * (gla - m);
            v = v * 0.999 + 0.0010000000000000009 * gla * gla;
            const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
            const double denom = sqrt(v) / sqrt(bc2) + 1e-8;
            const double la_new = la + (-(h.alpha_lr / bc1) * m) / denom;
            as[0] = la_new; as[1] = m; as[2] = v; as[3] = (double)t;
            ls[LOSS_ALPHA_VALUE] = (float)exp(la_new);
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// stage-boundary time stamp of agent 0 (profiling entry point spp_update_stage_profile; a null pointer in every other launch)
__device__ __forceinline__ void stage_mark(const Ctx& c, int g, int i) {
    if (c.a.timing && c.agent == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        c.a.timing[(size_t)g * kStageMarks + i] = t;
    }
}

template <int ALGO>
__device__ void update_step(const Ctx& c, int g) {
    const Layout& L = c.a.L;
    const Hyper& h = c.a.h;
    float* S = c.S;
    const int B = L.B;
    constexpr int ncrit = (ALGO == ALGO_SAC) ? 2 : 1;
    const int crit[2] = {NET_CRITIC_1, NET_CRITIC_2};
    const int tcrit[2] = {NET_CRITIC_1_TARG, NET_CRITIC_2_TARG};
    float* ls = c.sm.small + kLossBase;

    stage_mark(c, g, 0);
    // ---- bookkeeping
    if (threadIdx.x == 0) {
        if (ALGO == ALGO_SAC) c.sm.alpha = (float)exp(c.a.alpha_state[(size_t)c.agent * 4]);
        adam_begin(c, 0, h.actor_lr);
        for (int i = 0; i < ncrit; ++i) adam_begin(c, 1 + i, h.critic_lr);
        for (int k = 0; k < LOSS_COUNT; ++k) ls[k] = 0.f;
    }
    for (int i = threadIdx.x; i < 4 * 512; i += kThreads) __stcg(c.gvec(0) + i, 0.f);
    stage_gather(c, g);
    __syncthreads();
    stage_mark(c, g, 1);

    // ---- Phase A: Q target (sac_acm.py:43-56 / ddpg_acm.py:113-121)
    actor_forward<ALGO>(c, S + L.s.xn, ALGO == ALGO_SAC ? NET_ACTOR : NET_ACTOR_TARG);
    stage_mark(c, g, 2);
    if (ALGO == ALGO_SAC) stage_sample(c, g, 0); else stage_ddpg_post(c, 0);
    __syncthreads();
    stage_mark(c, g, 3);
    if (L.acm_critic) acm_forward(c, S + L.s.xcp + L.ldo, L.ldc);
    stage_mark(c, g, 4);
    critics_hidden(c, S + L.s.xcp, tcrit, ncrit, true);
    stage_mark(c, g, 5);
    stage_qtarget(c, tcrit, ncrit);
    __syncthreads();
    stage_mark(c, g, 6);

    // ---- Phase B: critic step(s) (sac_acm.py:117-131 / ddpg_acm.py:175-182), Polyak fused (sac.py:186-199)
    critics_hidden(c, S + L.s.xc, crit, ncrit, false);
    stage_mark(c, g, 7);
    stage_critic_head_bwd<0>(c, crit, tcrit, ncrit, ls + LOSS_CRITIC_1);
    __syncthreads();
    stage_mark(c, g, 8);
    for (int i = 0; i < ncrit; ++i) {   // dz1 = (dz2 W2) * relu'(hc1); column sums -> d b1
        const LayerDesc& l = L.critic.L[1];
        EpiMaskStore<MASK_RELU_BITS, true, false> epi{S + L.s.dz1[i], kHidden, S + L.s.hc1[i], kHidden, c.gvec(GV_CB1_0 + i),
                                                 reinterpret_cast<const unsigned long long*>(S + L.s.mk_hc1[i])};
        gemm_big<true>(c, S + L.s.dz2[i], kHidden, c.net(crit[i]) + l.off_w, l.ld, B, kHidden, kHidden, epi);
    }
    __syncthreads();
    stage_mark(c, g, 9);
    for (int i = 0; i < ncrit; ++i) {
        float* net = c.net(crit[i]); float* nm = c.net_m(crit[i]); float* nv = c.net_v(crit[i]); float* tn = c.net(tcrit[i]);
        const AdamScalars as = c.sm.adam[1 + i];
        {   // fc2.weight: dW2[m,n] = sum_b dz2[b,m] hc1[b,n]
            const LayerDesc& l = L.critic.L[1];
            EpiAdam epi{net + l.off_w, nm + l.off_w, nv + l.off_w, l.ld, net + l.off_wt, tn + l.off_wt, nullptr, l.ld_t, as, h.tau, h.one_minus_tau};
            gemm_big<false>(c, S + L.s.dz2[i], kHidden, S + L.s.hc1[i], kHidden, kHidden, kHidden, B, epi);
        }
        {   // fc1.weight: dW1[m,n] = sum_b dz1[b,m] xc[b,n]
            const LayerDesc& l = L.critic.L[0];
            EpiAdam epi{net + l.off_w, nm + l.off_w, nv + l.off_w, l.ld, net + l.off_wt, tn + l.off_wt, nullptr, l.ld_t, as, h.tau, h.one_minus_tau};
            gemm<NarrowTile, false>(S + L.s.dz1[i], kHidden, S + L.s.xc, L.ldc, kHidden, L.ldc, B, c.sm.gemm, epi);
            adam_vector(net + l.off_b, nm + l.off_b, nv + l.off_b, tn + l.off_b, c.gvec(GV_CB1_0 + i), kHidden, as, h.tau,
                        h.one_minus_tau, true);
        }
    }
    __syncthreads();
    stage_mark(c, g, 10);

    // ---- Phase C: policy step against the updated critic(s) (sac_acm.py:133-145,60-87 / ddpg_acm.py:125-145,187-192)
    actor_forward<ALGO>(c, S + L.s.xo, NET_ACTOR);
    stage_mark(c, g, 11);
    if (ALGO == ALGO_SAC) stage_sample(c, g, 1); else stage_ddpg_post(c, 1);
    __syncthreads();
    stage_mark(c, g, 12);
    if (L.acm_critic) acm_forward(c, S + L.s.xcp + L.ldo, L.ldc);
    stage_mark(c, g, 13);
    critics_hidden(c, S + L.s.xcp, crit, ncrit, true);
    stage_mark(c, g, 14);
    stage_policy_head_bwd(c, crit, ncrit, ls + LOSS_PI);
    __syncthreads();
    stage_mark(c, g, 15);
    for (int i = 0; i < ncrit; ++i) {
        const LayerDesc& l = L.critic.L[1];
        EpiMaskStore<MASK_RELU_BITS, false, false> epi{S + L.s.dz1[i], kHidden, S + L.s.hc1[i], kHidden, nullptr,
                                                  reinterpret_cast<const unsigned long long*>(S + L.s.mk_hc1[i])};
        gemm_big<true>(c, S + L.s.dz2[i], kHidden, c.net(crit[i]) + l.off_w, l.ld, B, kHidden, kHidden, epi);
    }
    __syncthreads();
    stage_mark(c, g, 16);
    {   // dxc = sum_i dz1[i] W1_i     [B x ldc]
        const LayerDesc& l = L.critic.L[0];
        EpiMaskStore<MASK_NONE, false, false> e0{S + L.s.dxc, L.ldc, nullptr, 0, nullptr};
        gemm<NarrowTile, true>(S + L.s.dz1[0], kHidden, c.net(crit[0]) + l.off_w, l.ld, B, L.ldc, kHidden, c.sm.gemm, e0);
        if (ncrit == 2) {
            EpiMaskStore<MASK_NONE, false, true> e1{S + L.s.dxc, L.ldc, nullptr, 0, nullptr};
            gemm<NarrowTile, true>(S + L.s.dz1[1], kHidden, c.net(crit[1]) + l.off_w, l.ld, B, L.ldc, kHidden, c.sm.gemm, e1);
        }
    }
    __syncthreads();
    stage_mark(c, g, 17);
    if (L.acm_critic) acm_backward_dx(c);
    stage_mark(c, g, 18);
    stage_actor_head_bwd<ALGO>(c);
    stage_mark(c, g, 19);
    {   // dza2 = (dml Wheads) * relu'(ha2); column sums -> d b2
        const LayerDesc& l = L.actor.L[2];
        EpiMaskStore<MASK_RELU_BITS, true, false> epi{S + L.s.dza2, kHidden, S + L.s.ha2, kHidden, c.gvec(GV_AB2),
                                                 reinterpret_cast<const unsigned long long*>(S + L.s.mk_ha2)};
        gemm_big<true>(c, S + L.s.dml, L.ldh, c.net(NET_ACTOR) + l.off_w, l.ld, B, kHidden, L.heads, epi);
    }
    __syncthreads();
    stage_mark(c, g, 20);
    {   // dza1 = (dza2 W2) * relu'(ha1); column sums -> d b1
        const LayerDesc& l = L.actor.L[1];
        EpiMaskStore<MASK_RELU_BITS, true, false> epi{S + L.s.dza1, kHidden, S + L.s.ha1, kHidden, c.gvec(GV_AB1),
                                                 reinterpret_cast<const unsigned long long*>(S + L.s.mk_ha1)};
        gemm_big<true>(c, S + L.s.dza2, kHidden, c.net(NET_ACTOR) + l.off_w, l.ld, B, kHidden, kHidden, epi);
    }
    __syncthreads();
    stage_mark(c, g, 21);
    {
        float* net = c.net(NET_ACTOR); float* nm = c.net_m(NET_ACTOR); float* nv = c.net_v(NET_ACTOR);
        float* tn = (ALGO == ALGO_DDPG) ? c.net(NET_ACTOR_TARG) : nullptr;
        const AdamScalars as = c.sm.adam[0];
        constexpr bool PK = (ALGO == ALGO_DDPG);     // DDPG.update_target_nets blends the actor too (ddpg.py:273-284)
        const LayerDesc& l2 = L.actor.L[2]; const LayerDesc& l1 = L.actor.L[1]; const LayerDesc& l0 = L.actor.L[0];
        {   // heads, computed transposed: tile[m = hidden unit, n = head row] = sum_b ha2[b, m] dml[b, n];
            // row copy = W^T (moments live in that layout), column copy = natural W
            EpiAdam epi{net + l2.off_wt, nm + l2.off_wt, nv + l2.off_wt, l2.ld_t, net + l2.off_w, nullptr,
                        PK ? tn + l2.off_wt : nullptr, l2.ld, as, h.tau, h.one_minus_tau};
            gemm<NarrowTile, false>(S + L.s.ha2, kHidden, S + L.s.dml, L.ldh, kHidden, L.heads, B, c.sm.gemm, epi);
        }
        {
            EpiAdam epi{net + l1.off_w, nm + l1.off_w, nv + l1.off_w, l1.ld, net + l1.off_wt, PK ? tn + l1.off_wt : nullptr,
                        nullptr, l1.ld_t, as, h.tau, h.one_minus_tau};
            gemm_big<false>(c, S + L.s.dza2, kHidden, S + L.s.ha1, kHidden, kHidden, kHidden, B, epi);
        }
        {
            EpiAdam epi{net + l0.off_w, nm + l0.off_w, nv + l0.off_w, l0.ld, net + l0.off_wt, PK ? tn + l0.off_wt : nullptr,
                        nullptr, l0.ld_t, as, h.tau, h.one_minus_tau};
            gemm<NarrowTile, false>(S + L.s.dza1, kHidden, S + L.s.xo, L.ldo, kHidden, L.ldo, B, c.sm.gemm, epi);
        }
        adam_vector(net + l2.off_b, nm + l2.off_b, nv + l2.off_b, PK ? tn + l2.off_b : nullptr, c.gvec(GV_MISC), L.heads, as, h.tau, h.one_minus_tau, true);
        adam_vector(net + l1.off_b, nm + l1.off_b, nv + l1.off_b, PK ? tn + l1.off_b : nullptr, c.gvec(GV_AB2), kHidden, as, h.tau, h.one_minus_tau, true);
        adam_vector(net + l0.off_b, nm + l0.off_b, nv + l0.off_b, PK ? tn + l0.off_b : nullptr, c.gvec(GV_AB1), kHidden, as, h.tau, h.one_minus_tau, true);
    }
    __syncthreads();
    stage_mark(c, g, 22);
    if (threadIdx.x < LOSS_COUNT && c.a.losses)
        c.a.losses[((size_t)c.agent * c.a.G + g) * LOSS_COUNT + threadIdx.x] = ls[threadIdx.x];
    __syncthreads();
    stage_mark(c, g, 23);
}

template <int ALGO>
__global__ void __launch_bounds__(kThreads, 1) update_burst_kernel(const __grid_constant__ UpdateArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = smem_struct(smem_raw);
    UmmaCtx* um = a.use_umma ? umma_setup(sm, a.use_umma) : nullptr;
    for (int agent = blockIdx.x; agent < a.population; agent += gridDim.x) {
        Ctx c(a, agent, sm, um);
        for (int g = 0; g < a.G; ++g) update_step<ALGO>(c, g);
    }
    if (um) umma_teardown(um);
}

// More agents than CTAs (config 3: 256 agents on 148 SMs): whole agents per CTA (the kernel above) take ceil(P / grid) * G step
// times with the last wave partly empty (2 x 50 steps for 256 agents).  Here the P * G (step, agent) items run step-major,
// interleaved over the CTAs: CTA k takes items k, k + grid, ...  -> ceil(P * G / grid) step times (87 instead of 100).  Nothing of an
// agent lives in a CTA between steps (weights, moments, scratch and step counters are all in global memory), so an agent may change
// CTA every step; item (g, agent) only has to wait for (g - 1, agent), which another CTA finished one or two of its own items
// earlier: a release / acquire pair on the agent's progress word (a.progress, zeroed by the host before the launch).  Items are taken
// in increasing order and only ever wait for smaller ones, and the grid (<= one CTA per SM) is co-resident, so the smallest
// unfinished item can always run.  The wait is bounded all the same: a kernel that can never finish must not hang the GPU.
// A separate kernel so that the one-agent-per-CTA kernel (the headline configuration) keeps its code unchanged.
template <int ALGO>
__global__ void __launch_bounds__(kThreads, 1) update_burst_interleaved_kernel(const __grid_constant__ UpdateArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = smem_struct(smem_raw);
    UmmaCtx* um = a.use_umma ? umma_setup(sm, a.use_umma) : nullptr;
    // an item is kItemSteps consecutive steps of one agent: the step it waits for then lies 2 P / grid > 2 items back (no CTA ever
    // waits in steady state) at the price of a slightly coarser tail (88 instead of 87 step times for 256 agents x 50 steps)
    constexpr int kItemSteps = 2;
    const int chunks = (a.G + kItemSteps - 1) / kItemSteps;
    const int total = a.population * chunks;      // < 2^31 (checked by the host)
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int q = w / a.population, agent = w - q * a.population;
        const int g0 = q * kItemSteps, g1 = min(a.G, g0 + kItemSteps);
        if (g0 > 0) {
            if (threadIdx.x == 0) {
                unsigned int done, spins = 0;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(a.progress + agent) : "memory");
                    if (done >= (unsigned int)g0) break;
                    if (++spins > (1u << 24)) __trap();      // ~10 s: the step this item waits for is lost
                    __nanosleep(64);
                }
                __threadfence();
            }
            __syncthreads();
        }
        Ctx c(a, agent, sm, um);
        for (int g = g0; g < g1; ++g) update_step<ALGO>(c, g);      // each ends with a CTA barrier behind every global write of the step
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.progress + agent), "r"((unsigned int)g1) : "memory");
        }
    }
    if (um) umma_teardown(um);
}

template <class K>
static cudaError_t launch_burst(K kernel, const UpdateArgs& a, int grid, cudaStream_t stream) {
    const size_t smem = kSmemLaunchBytes;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_update_burst(const UpdateArgs& a, int grid, cudaStream_t stream) {
    if (a.progress)
        return a.L.algo == ALGO_SAC ? launch_burst(update_burst_interleaved_kernel<ALGO_SAC>, a, grid, stream)
                                    : launch_burst(update_burst_interleaved_kernel<ALGO_DDPG>, a, grid, stream);
    return a.L.algo == ALGO_SAC ? launch_burst(update_burst_kernel<ALGO_SAC>, a, grid, stream)
                                : launch_burst(update_burst_kernel<ALGO_DDPG>, a, grid, stream);
}

}  // namespace spp
