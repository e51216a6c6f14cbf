// ACM (inverse-dynamics model) regression bursts: n consecutive minibatch steps per agent in one launch.
// Reference: AcMTrainer.update_acm_batches (rltoolkit/acm/acm.py:356-372) = n x [rbuffer_sample_acm
// (rltoolkit/buffer/replay_buffer.py:404-430) -> acm_cat (acm.py:260-264) -> batch_update (acm.py:246-258)]:
// MSE(acm(cat[obs, next_obs]), actions_acm), Adam(acm_lr).  No LR-scheduler step on this path (quirk 20).
#include "update_kernel.cuh"

namespace spp {

__device__ inline void acm_gather(const Ctx& c, int g) {
    const Layout& L = c.a.L;
    const int B = L.B, ob = L.ob, ldo = L.ldo, lane = lane_id();
    float* xm = c.S + L.s.xm; float* ya = c.S + L.s.ya;
    const bool from_ring = (c.a.batch.acm_x == nullptr);
    for (int r = warp_id(); r < B; r += kWarps) {
        if (from_ring) {
            const RingPtrs& R = c.a.ring;
            int64_t i;
            if (c.a.idx) {
                i = c.a.idx[((size_t)c.agent * c.a.G + g) * (c.a.batch_row_stride > 0 ? c.a.batch_row_stride : B) + r];
            } else {
                const uint4 x = Philox::gen(c.a.seed ^ 0x51ED270B0B5ull, ((uint64_t)c.agent << 32) | (uint32_t)g,
                                            (c.a.seq << 20) | (uint32_t)r);
                i = (int64_t)__umul64hi(((uint64_t)x.x << 32) | x.y, (uint64_t)c.a.ring_len[c.agent]);
            }
            const size_t base = (size_t)c.agent * R.S;
            const float* po = R.obs + (base + R.oidx[base + i]) * ldo;
            const float* pn = R.obs + (base + R.nidx[base + i]) * ldo;
            for (int j = lane; j < ob; j += 32) { xm[r * L.ldm + j] = po[j]; xm[r * L.ldm + ldo + j] = pn[j]; }
            for (int j = lane; j < L.ac; j += 32) ya[r * L.lda + j] = R.aacm[(base + i) * L.lda + j];
        } else {
            const int stride = c.a.batch_row_stride > 0 ? c.a.batch_row_stride : B;
            const size_t row = ((size_t)c.agent * c.a.G + g) * stride + r;
            const float* px = c.a.batch.acm_x + row * 2 * ob;
            for (int j = lane; j < ob; j += 32) { xm[r * L.ldm + j] = px[j]; xm[r * L.ldm + ldo + j] = px[ob + j]; }
            for (int j = lane; j < L.ac; j += 32) ya[r * L.lda + j] = c.a.batch.acm_y[row * L.ac + j];
        }
    }
}

// gvec slots used here: 0 = d b2 (pre-activation of hidden 2), 1 = d b1, 2 = d b21 (BasicAcM skip), 3 = d b3 / gains
__device__ void acm_train_step(const Ctx& c, int g) {
    const Layout& L = c.a.L;
    const Hyper& h = c.a.h;
    float* S = c.S;
    const int B = L.B, ac = L.ac;
    const bool basic = (L.acm_kind == ACM_BASIC);
    float* acm = c.net(NET_ACM); float* nm = c.net_m(NET_ACM); float* nv = c.net_v(NET_ACM);
    const bool eval = c.a.acm_eval != 0;
    if (threadIdx.x == 0 && !eval) adam_begin(c, 3, h.acm_lr);
    for (int i = threadIdx.x; i < 4 * 512; i += kThreads) __stcg(c.gvec(0) + i, 0.f);
    acm_gather(c, g);
    __syncthreads();
    // ---- forward (keeps hm1, hm2, tm3; the scaled prediction goes to the action block of xcp)
    acm_forward(c, S + L.s.pa, L.lda);
    // ---- loss and d3 = (2 / (B ac)) (pred - y) * scale * (1 - t3^2); column sums -> d b3 (and d t1)
    const float* scale = basic ? acm + L.acm.L[4].off_w + 4 : c.a.acm_lim;
    const float norm = (float)(2.0 / ((double)B * (double)ac));
    float lsum = 0.f;
    for (int e = threadIdx.x; e < B * ac; e += kThreads) {
        const int r = e / ac, j = e % ac;
        const float t3 = S[L.s.tm3 + r * L.lda + j];
        const float pred = S[L.s.pa + r * L.lda + j];
        const float diff = __fsub_rn(pred, S[L.s.ya + r * L.lda + j]);
        lsum = fmaf(diff, diff, lsum);
        const float da = __fmul_rn(norm, diff);
        S[L.s.dm3 + r * L.lda + j] = __fmul_rn(__fmul_rn(da, scale[j]), __fsub_rn(1.f, __fmul_rn(t3, t3)));
        if (basic) S[L.s.pa + r * L.lda + j] = __fmul_rn(da, t3);      // d t1 contribution (prediction no longer needed)
    }
    const float ltot = block_sum(lsum, c.sm.small);
    __syncthreads();
    if (threadIdx.x == 0 && c.a.losses) c.a.losses[(size_t)c.agent * c.a.G + g] = ltot / ((float)B * (float)ac);
    if (eval) return;      // AcMTrainer.calculate_validation_loss (rltoolkit/acm/acm.py:329-343): no_grad forward + loss
    if (threadIdx.x < ac) {      // deterministic column sums over the batch
        float s3 = 0.f, st1 = 0.f;
        for (int r = 0; r < B; ++r) {
            s3 += S[L.s.dm3 + r * L.lda + threadIdx.x];
            if (basic) st1 += S[L.s.pa + r * L.lda + threadIdx.x];
        }
        c.sm.vecs[threadIdx.x] = s3;                 // d b3
        if (basic) c.sm.vecs[64 + 4 + threadIdx.x] = st1;   // gains gradient vector: [t, 0, 0, 0, t1...]
    }
    __syncthreads();
    // ---- backward
    {   // d2 = (d3 W3) * (1 - hm2^2), column sums -> d b2
        const LayerDesc& l = L.acm.L[2];
        EpiMaskStore<MASK_TANH, true, false> epi{S + L.s.dm2, L.ldm2, S + L.s.hm2, L.ldm2, c.gvec(0)};
        gemm<NarrowTile, true>(S + L.s.dm3, L.lda, acm + l.off_w, l.ld, B, L.hm2, ac, c.sm.gemm, epi);
    }
    __syncthreads();
    if (basic) {   // d t = sum d2 * s ; ds = d2 * t
        const float t = acm[L.acm.L[4].off_w];
        float ts = 0.f;
        for (int e = threadIdx.x; e < B * L.hm2; e += kThreads) {
            const int r = e / L.hm2, j = e % L.hm2;
            const float d2 = S[L.s.dm2 + r * L.ldm2 + j];
            ts = fmaf(d2, S[L.s.hms + r * L.ldm2 + j], ts);
            S[L.s.dms + r * L.ldm2 + j] = __fmul_rn(d2, t);
        }
        const float tt = block_sum(ts, c.sm.small);
        if (threadIdx.x == 0) { c.sm.vecs[64] = tt; c.sm.vecs[65] = 0.f; c.sm.vecs[66] = 0.f; c.sm.vecs[67] = 0.f; }
        __syncthreads();
        for (int j = threadIdx.x; j < L.hm2; j += kThreads) {   // d b21 = column sums of ds
            float s = 0.f;
            for (int r = 0; r < B; ++r) s += S[L.s.dms + r * L.ldm2 + j];
            __stcg(c.gvec(2) + j, s);
        }
    }
    {   // d1 = (d2 W2) * (1 - hm1^2), column sums -> d b1
        const LayerDesc& l = L.acm.L[1];
        EpiMaskStore<MASK_TANH, true, false> epi{S + L.s.dm1, L.ldm1, S + L.s.hm1, L.ldm1, c.gvec(1)};
        gemm<NarrowTile, true>(S + L.s.dm2, L.ldm2, acm + l.off_w, l.ld, B, L.hm1, L.hm2, c.sm.gemm, epi);
    }
    __syncthreads();
    // ---- weight gradients + Adam (row copy = natural W, column copy = W^T)
    const AdamScalars as = c.sm.adam[3];
    auto adam_layer = [&](int li, const float* dY, int ldy, const float* X, int ldx) {
        const LayerDesc& l = L.acm.L[li];
        EpiAdam epi{acm + l.off_w, nm + l.off_w, nv + l.off_w, l.ld, acm + l.off_wt, nullptr, nullptr, l.ld_t, as, 0.f, 1.f};
        gemm<NarrowTile, false>(dY, ldy, X, ldx, l.rows, l.ld, B, c.sm.gemm, epi);
    };
    adam_layer(2, S + L.s.dm3, L.lda, S + L.s.hm2, L.ldm2);
    adam_layer(1, S + L.s.dm2, L.ldm2, S + L.s.hm1, L.ldm1);
    adam_layer(0, S + L.s.dm1, L.ldm1, S + L.s.xm, L.ldm);
    if (basic) adam_layer(3, S + L.s.dms, L.ldm2, S + L.s.xm, L.ldm);
    {
        const LayerDesc& l2 = L.acm.L[2]; const LayerDesc& l1 = L.acm.L[1]; const LayerDesc& l0 = L.acm.L[0];
        adam_vector(acm + l2.off_b, nm + l2.off_b, nv + l2.off_b, nullptr, c.sm.vecs, ac, as, 0.f, 1.f, false);
        adam_vector(acm + l1.off_b, nm + l1.off_b, nv + l1.off_b, nullptr, c.gvec(0), L.hm2, as, 0.f, 1.f, true);
        adam_vector(acm + l0.off_b, nm + l0.off_b, nv + l0.off_b, nullptr, c.gvec(1), L.hm1, as, 0.f, 1.f, true);
        if (basic) {
            const LayerDesc& l3 = L.acm.L[3]; const LayerDesc& l4 = L.acm.L[4];
            adam_vector(acm + l3.off_b, nm + l3.off_b, nv + l3.off_b, nullptr, c.gvec(2), L.hm2, as, 0.f, 1.f, true);
            adam_vector(acm + l4.off_w, nm + l4.off_w, nv + l4.off_w, nullptr, c.sm.vecs + 64, 4 + ac, as, 0.f, 1.f, false);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) acm_train_kernel(const __grid_constant__ UpdateArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = smem_struct(smem_raw);
    for (int agent = blockIdx.x; agent < a.population; agent += gridDim.x) {
        Ctx c(a, agent, sm);
        const int full = (a.acm_last_rows > 0) ? a.G - 1 : a.G;
        for (int g = 0; g < full; ++g) acm_train_step(c, g);
        if (a.acm_last_rows > 0) {       // DataLoader keeps the partial last minibatch (drop_last = False)
            UpdateArgs tail = a;
            tail.L.B = a.acm_last_rows;
            tail.batch_row_stride = a.L.B;
            Ctx ct(tail, agent, sm);
            acm_train_step(ct, a.G - 1);
        }
    }
}

cudaError_t launch_acm_train(const UpdateArgs& a, int grid, cudaStream_t stream) {
    const size_t smem = kSmemLaunchBytes;
    cudaError_t e = cudaFuncSetAttribute(acm_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    acm_train_kernel<<<grid, kThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace spp
