// Vectorised rollout step: normalise -> actor -> exploration noise / clip -> denormalise -> ACM -> ring write,
// for E environments per agent that share the agent's weights (the reference steps ONE environment at a time).
// Reference: body of DDPG.collect_batch_and_train (rltoolkit/algorithms/ddpg/ddpg.py:202-223) with
// AcMOffPolicy.initial_act (rltoolkit/acm/off_policy/off_policy.py:50-54), DDPG_AcM.noise_action
// (rltoolkit/acm/off_policy/ddpg_acm.py:40-50), AcMOffPolicy.process_action (off_policy.py:89-106) and the ring
// writes of rltoolkit/buffer/replay_buffer.py:56-75,133-137,332-333.  MuJoCo is not available: the device-resident
// form advances a synthetic environment (obs' = 0.98 obs + 0.1 tanh(a-mix) + 0.02 N(0,1)) so that everything the
// reference does per frame except the simulator itself is on the device.
#define SPP_UMMA_WRAPPER_INLINE 1      // see gemm_umma.cuh
#include "update_kernel.cuh"

namespace spp {

// obs -> (normalised) actor input xo
__device__ inline void rollout_load(const Ctx& c, const RolloutArgs& r, int step) {
    const Layout& L = c.a.L;
    const int E = L.B, ob = L.ob, ldo = L.ldo;
    float* xo = c.S + L.s.xo;
    const float* nsub = c.normv(NORM_NSUB); const float* ndiv = c.normv(NORM_NDIV);
    for (int e = threadIdx.x; e < E * ob; e += kThreads) {
        const int row = e / ob, j = e % ob;
        float v = r.env_state ? r.env_state[((size_t)c.agent * E + row) * ldo + j] : r.in_obs[((size_t)c.agent * E + row) * ob + j];
        if (r.obs_norm) {
            v = __fdiv_rn(__fsub_rn(v, nsub[j]), ndiv[j]);
            if (c.a.h.norm_clamp) v = fminf(fmaxf(v, -10.f), 10.f);
        }
        xo[row * ldo + j] = v;
    }
}

// heads -> state target (noise, clip, denormalise); builds the ACM input [x | target]
template <int ALGO>
__device__ inline void rollout_post(const Ctx& c, const RolloutArgs& r, int step) {
    const Layout& L = c.a.L;
    const int E = L.B, ob = L.ob, ldo = L.ldo;
    const float* xo = c.S + L.s.xo; const float* ml = c.S + L.s.ml; float* xm = c.S + L.s.xm;
    const float* lim = c.normv(NORM_LIM); const float* doff = c.normv(NORM_DOFF); const float* dsc = c.normv(NORM_DSCALE);
    for (int e = threadIdx.x; e < E * ob; e += kThreads) {
        const int row = e / ob, j = e % ob;
        const size_t gi = ((size_t)c.agent * E + row) * ob + j;
        float nz, ez = 0.f;
        if (r.in_noise) nz = r.in_noise[gi];
        else {
            const uint4 w = Philox::gen(c.a.seed ^ 0x7001ull, ((uint64_t)c.agent << 32) | (uint32_t)step, (c.a.seq << 24) | (uint32_t)e);
            nz = normal_from_bits(w.x, w.y);
            ez = normal_from_bits(w.z, w.w);
        }
        float z;
        if (r.random_phase == 2) {
            z = nz;                                                      // caller supplies the state target (process_action only)
        } else if (r.random_phase) {
            z = __fmul_rn(lim[j], nz);                                   // actor_ac_lim * randn (off_policy.py:51)
        } else {
            if (ALGO == ALGO_SAC) {
                if (r.in_eps) ez = r.in_eps[gi];
                const float mu = ml[row * L.ldh + j];
                const float ls = fminf(fmaxf(ml[row * L.ldh + ob + j], -20.f), 2.f);
                const float u = __fadd_rn(mu, __fmul_rn(ez, expf(ls)));
                z = __fmul_rn(tanhf(u), lim[j]);
            } else {
                z = ml[row * L.ldh + j];                                 // tanh(fc3) * lim from the head epilogue
            }
            z = __fadd_rn(z, __fmul_rn(__fmul_rn(r.act_noise, nz), lim[j]));   // action += noise * actor_ac_lim
            const float hi = __fmul_rn(1.1f, lim[j]);
            z = fminf(fmaxf(z, -hi), hi);                                // np.clip(action, -1.1 lim, 1.1 lim)
        }
        if (r.denormalize_out) z = __fadd_rn(doff[j], __fmul_rn(z, dsc[j]));
        xm[row * L.ldm + j] = xo[row * ldo + j];
        xm[row * L.ldm + ldo + j] = z;
        if (r.out_target) r.out_target[gi] = z;
    }
}

// environment step (synthetic) + ring writes.  Row mapping: env e at step t owns obs row obs_cur + t*E + e and
// timestep row ts_cur + t*E + e; its next observation is the row E further on, so every step adds one obs row
// and one timestep row per environment, like add_obs / add_timestep do.
__device__ inline void rollout_env_and_ring(const Ctx& c, const RolloutArgs& r, int step) {
    const Layout& L = c.a.L;
    const int E = L.B, ob = L.ob, ac = L.ac, ldo = L.ldo;
    const float* xm = c.S + L.s.xm; const float* pa = c.S + L.s.pa;
    const int64_t S = c.a.ring.S;
    const size_t base = (size_t)c.agent * S;
    const int64_t o0 = r.obs_cur[c.agent] + (int64_t)step * E, t0 = r.ts_cur[c.agent] + (int64_t)step * E;
    // a lane group of gs lanes (smallest power of two >= ob, at least 4) per environment: 32 / gs environments per warp at a time;
    // the group-wide xor tree gives the same sum as a 32-lane tree over zero-padded lanes
    int gs = 32;
    while (gs > 4 && (gs >> 1) >= ob) gs >>= 1;
    const int rpw = 32 / gs, grp = lane_id() / gs, lane = lane_id() % gs;
    for (int rb = warp_id() * rpw; rb < E; rb += kWarps * rpw) {
        const int row = rb + grp;
        const bool valid = row < E;
        // cursors < S, row < E and (steps + 1) E <= S: at most two wraps -- conditional subtraction instead of a 64-bit modulo
        int64_t orow = o0 + row, nrow = o0 + E + row, trow = t0 + row;
        orow -= orow >= S ? S : 0; orow -= orow >= S ? S : 0;
        nrow -= nrow >= S ? S : 0; nrow -= nrow >= S ? S : 0;
        trow -= trow >= S ? S : 0; trow -= trow >= S ? S : 0;
        float* st = r.env_state + ((size_t)c.agent * E + (valid ? row : 0)) * ldo;
        float mix = 0.f;
        if (valid)
            for (int j = lane; j < ac; j += gs) {
                const float a = pa[row * L.lda + j];
                r.w_aacm[(base + trow) * L.lda + j] = a;
                mix += a * (0.3f + 0.1f * (float)j);
            }
        for (int o = gs >> 1; o > 0; o >>= 1) mix += __shfl_xor_sync(0xffffffffu, mix, o);
        mix = tanhf(mix);
        if (!valid) continue;
        float rew = 0.f;
        for (int j = lane; j < ob; j += gs) {
            const uint4 w = Philox::gen(c.a.seed ^ 0xE9Full, ((uint64_t)c.agent << 32) | (uint32_t)step, (c.a.seq << 24) | (uint32_t)(row * ob + j));
            const float o = st[j];
            const float nx = 0.98f * o + 0.1f * mix * (1.f - 0.01f * (float)j) + 0.02f * normal_from_bits(w.x, w.y);
            if (step == 0) r.w_obs[(base + orow) * ldo + j] = o;          // the first observation row of this launch
            r.w_obs[(base + nrow) * ldo + j] = nx;
            if (r.w_act) r.w_act[(base + trow) * ldo + j] = xm[row * L.ldm + ldo + j];
            st[j] = nx;
            if (j == 0) rew = nx;
        }
        if (lane == 0) {
            const uint4 w = Philox::gen(c.a.seed ^ 0xD0Eull, ((uint64_t)c.agent << 32) | (uint32_t)step, (c.a.seq << 24) | (uint32_t)row);
            r.w_oidx[base + trow] = (int32_t)orow;
            r.w_nidx[base + trow] = (int32_t)nrow;
            r.w_rew[base + trow] = rew;
            r.w_done[base + trow] = (w.x < 4294967u) ? 1 : 0;
            r.w_end[base + trow] = 0;
        }
    }
}

template <int ALGO>
__global__ void __launch_bounds__(kThreads, 1) rollout_kernel(const __grid_constant__ RolloutArgs r) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = smem_struct(smem_raw);
    const UpdateArgs& a = r.u;
    UmmaCtx* um = a.use_umma ? umma_setup(sm, a.use_umma) : nullptr;
    for (int agent = blockIdx.x; agent < a.population; agent += gridDim.x) {
        Ctx c(a, agent, sm, um);
        const Layout& L = a.L;
        for (int step = 0; step < r.steps; ++step) {
            rollout_load(c, r, step);
            __syncthreads();
            if (!r.random_phase) actor_forward<ALGO>(c, c.S + L.s.xo, NET_ACTOR);
            rollout_post<ALGO>(c, r, step);
            __syncthreads();
            acm_forward(c, c.S + L.s.pa, L.lda);
            if (r.out_action) {
                for (int e = threadIdx.x; e < L.B * L.ac; e += kThreads)
                    r.out_action[((size_t)agent * L.B + e / L.ac) * L.ac + e % L.ac] = c.S[L.s.pa + (e / L.ac) * L.lda + e % L.ac];
            }
            if (r.env_state) rollout_env_and_ring(c, r, step);
            __syncthreads();
        }
    }
    if (um) umma_teardown(um);
}

cudaError_t launch_rollout(const RolloutArgs& r, int grid, cudaStream_t stream) {
    const size_t smem = kSmemLaunchBytes;
    cudaError_t e;
    if (r.u.L.algo == ALGO_SAC) {
        e = cudaFuncSetAttribute(rollout_kernel<ALGO_SAC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rollout_kernel<ALGO_SAC><<<grid, kThreads, smem, stream>>>(r);
    } else {
        e = cudaFuncSetAttribute(rollout_kernel<ALGO_DDPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rollout_kernel<ALGO_DDPG><<<grid, kThreads, smem, stream>>>(r);
    }
    return cudaGetLastError();
}

}  // namespace spp
