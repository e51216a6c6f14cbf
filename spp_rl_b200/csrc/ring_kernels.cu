// Replay-ring kernels: two-level minibatch gather, row writes, synthetic prefill.
// Reference semantics: rltoolkit/buffer/replay_buffer.py:233-261,385-398 (sample), :56-75,133-137,332-333 (adds).
#include "common.cuh"
#include "ring_kernels.h"
#include "ppo_rollout.h"

namespace spp {

// One warp per sampled row: obs = ring_obs[oidx[i]], next_obs = ring_obs[nidx[i]], plus the per-timestep
// columns.  Rows are padded to 16 B in the ring; outputs are dense (the reference's tensor shapes).
__global__ void ring_gather_kernel(RingView R, int agent, const int64_t* __restrict__ idx, int n, GatherOut o) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const size_t base = (size_t)agent * R.S;
    for (int r = w; r < n; r += nw) {
        const int64_t i = idx[r];
        const float* po = R.obs + (base + R.oidx[base + i]) * R.ldo;
        const float* pn = R.obs + (base + R.nidx[base + i]) * R.ldo;
        for (int j = lane; j < R.ob; j += 32) {
            o.obs[(size_t)r * R.ob + j] = po[j];
            o.nobs[(size_t)r * R.ob + j] = pn[j];
            if (o.act && R.act) o.act[(size_t)r * R.ob + j] = R.act[(base + i) * R.ldo + j];
        }
        for (int j = lane; j < R.ac; j += 32) o.aacm[(size_t)r * R.ac + j] = R.aacm[(base + i) * R.lda + j];
        if (lane == 0) { o.rew[r] = R.rew[base + i]; o.done[r] = (int8_t)R.done[base + i]; }
    }
}

// Population-wide gather benchmark: rows for (agent, batch, row) with device-drawn indices into dense minibatches [P][nb][B][...] --
// the HBM side of sample_batch with nothing else attached.  Two phases per warp and 32 rows, like stage_gather of the fused update
// kernel: (1) every lane draws ONE row index and loads its obs / next-obs slots (32 independent index -> slot chains in flight per
// warp), (2) groups of four lanes copy one row each (float4 per lane), eight rows at a time with every load issued before the first
// store.  A warp per row (round 1) kept three lanes busy and one dependent chain in flight: 363 GB/s algorithmic.
__global__ void __launch_bounds__(256) ring_gather_bench_kernel(RingView R, int P, int nb, int B, const int64_t* __restrict__ len, uint64_t seed,
                                                                float* __restrict__ out_obs, float* __restrict__ out_nobs,
                                                                float* __restrict__ out_aacm, float* __restrict__ out_rew,
                                                                uint8_t* __restrict__ out_done) {
    const int lane = threadIdx.x & 31, sub = lane >> 2, l4 = lane & 3;
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t total = (size_t)P * nb * B, per_agent = (size_t)nb * B;
    const int ldo4 = R.ldo / 4, lda4 = R.lda / 4;
    for (size_t t0 = w * 32; t0 < total; t0 += nw * 32) {
        // phase 1: lane -> row t0 + lane
        const size_t t = t0 + lane;
        const bool ok = t < total;
        const int agent = ok ? (int)(t / per_agent) : 0;
        const size_t base = (size_t)agent * R.S;
        int64_t i = 0; int32_t oi = 0, ni = 0; float rew = 0.f; uint8_t dn = 0;
        if (ok) {
            const uint4 x = Philox::gen(seed, (uint64_t)agent, (uint64_t)(t % per_agent));
            i = (int64_t)__umul64hi(((uint64_t)x.x << 32) | x.y, (uint64_t)len[agent]);
            oi = __ldg(R.oidx + base + i); ni = __ldg(R.nidx + base + i);
            rew = __ldg(R.rew + base + i); dn = __ldg(R.done + base + i);
        }
        if (ok) { out_rew[t] = rew; out_done[t] = dn; }
        // phase 2: four lanes per row, eight rows per step
#pragma unroll
        for (int step = 0; step < 4; ++step) {
            const int src = step * 8 + sub;
            const size_t tr = t0 + src;
            const int64_t ir = __shfl_sync(0xffffffffu, i, src);
            const int32_t oir = __shfl_sync(0xffffffffu, oi, src), nir = __shfl_sync(0xffffffffu, ni, src);
            const int ar = __shfl_sync(0xffffffffu, agent, src);
            if (tr >= total) continue;
            const size_t br = (size_t)ar * R.S;
            const float4* po = reinterpret_cast<const float4*>(R.obs + (br + oir) * R.ldo);
            const float4* pn = reinterpret_cast<const float4*>(R.obs + (br + nir) * R.ldo);
            float4* qo = reinterpret_cast<float4*>(out_obs + tr * R.ldo);
            float4* qn = reinterpret_cast<float4*>(out_nobs + tr * R.ldo);
            if (ldo4 <= 4) {      // rows of up to 16 floats: one float4 per lane, loads first
                float4 vo = make_float4(0.f, 0.f, 0.f, 0.f), vn = vo, va = vo;
                if (l4 < ldo4) { vo = __ldg(po + l4); vn = __ldg(pn + l4); }
                if (l4 < lda4) va = __ldg(reinterpret_cast<const float4*>(R.aacm + (br + ir) * R.lda) + l4);
                if (l4 < ldo4) { qo[l4] = vo; qn[l4] = vn; }
                if (l4 < lda4) reinterpret_cast<float4*>(out_aacm + tr * R.lda)[l4] = va;
            } else {
                for (int j = l4; j < ldo4; j += 4) { const float4 vo = __ldg(po + j), vn = __ldg(pn + j); qo[j] = vo; qn[j] = vn; }
                for (int j = l4; j < lda4; j += 4)
                    reinterpret_cast<float4*>(out_aacm + tr * R.lda)[j] = __ldg(reinterpret_cast<const float4*>(R.aacm + (br + ir) * R.lda) + j);
            }
        }
    }
}

// Synthetic prefill: n transitions per agent in episodes of T steps (T+1 obs rows per episode).
__global__ void ring_fill_kernel(RingView R, float* obs, int32_t* oidx, int32_t* nidx, float* act, float* rew,
                                 uint8_t* done, uint8_t* end, float* aacm, int P, int64_t n, int T, uint64_t seed,
                                 const float* __restrict__ norm, int norm_stride) {
    const size_t total = (size_t)P * n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int agent = (int)(t / n);
        const int64_t i = (int64_t)(t % n);
        const int64_t ep = i / T, st = i % T;
        const int64_t orow = ep * (T + 1) + st;
        const size_t base = (size_t)agent * R.S;
        oidx[base + i] = (int32_t)orow;
        nidx[base + i] = (int32_t)(orow + 1);
        const float* doff = norm + (size_t)agent * norm_stride;          // NORM_DOFF
        const float* dsc = doff + R.ldo;                                 // NORM_DSCALE
        for (int j = 0; j < R.ob; j += 2) {
            const uint4 x = Philox::gen(seed, (uint64_t)agent, (uint64_t)i * 128 + j);
            const float u0 = (float)(x.x >> 8) * (1.0f / 16777216.0f), u1 = (float)(x.y >> 8) * (1.0f / 16777216.0f);
            const float o0 = doff[j] + (2.f * u0 - 1.f) * dsc[j];
            obs[(base + orow) * R.ldo + j] = o0;
            if (st == T - 1 || i == n - 1) obs[(base + orow + 1) * R.ldo + j] = o0 + 0.02f * normal_from_bits(x.z, x.w);
            if (act) act[(base + i) * R.ldo + j] = 2.f * u1 - 1.f;
            if (j + 1 < R.ob) {
                const float o1 = doff[j + 1] + (2.f * u1 - 1.f) * dsc[j + 1];
                obs[(base + orow) * R.ldo + j + 1] = o1;
                if (st == T - 1 || i == n - 1) obs[(base + orow + 1) * R.ldo + j + 1] = o1 + 0.02f * normal_from_bits(x.w, x.z);
                if (act) act[(base + i) * R.ldo + j + 1] = 2.f * u0 - 1.f;
            }
        }
        const uint4 y = Philox::gen(seed ^ 0xABCDEF12345ull, (uint64_t)agent, (uint64_t)i);
        rew[base + i] = normal_from_bits(y.x, y.y);
        done[base + i] = (y.z < 4294967u) ? 1 : 0;      // ~1e-3
        end[base + i] = (st == T - 1) ? 1 : 0;
        for (int j = 0; j < R.ac; ++j) {
            const uint4 z = Philox::gen(seed ^ 0x5555AAAA5555ull, (uint64_t)agent, (uint64_t)i * 16 + j);
            aacm[(base + i) * R.lda + j] = tanhf(normal_from_bits(z.x, z.y));
        }
    }
}

// ReplayBufferAcM.add_buffer on device data (rltoolkit/buffer/replay_buffer.py:284-297): the host walked the cursor state machine
// (integers only) and left, per ring slot, where its final content comes from; every slot is written by exactly one thread group.
__global__ void ring_add_store_kernel(float* r_obs, int32_t* r_oidx, int32_t* r_nidx, float* r_aacm, float* r_rew, uint8_t* r_done,
                                      uint8_t* r_end, int64_t S, int ob, int ac, int ldo, int lda, const int64_t* __restrict__ obs_src,
                                      const int64_t* __restrict__ ts_src, const int32_t* __restrict__ ts_oidx,
                                      const int32_t* __restrict__ ts_nidx, const float* __restrict__ raw_obs,
                                      const float* __restrict__ raw_next, const float* __restrict__ aacm) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = obs_src[s];
        if (o >= 0) {
            const bool nxt = (o >> 62) & 1;
            const int64_t row = o & ((1ll << 62) - 1);
            const float* src = (nxt ? raw_next : raw_obs) + row * ldo;
            for (int j = 0; j < ob; ++j) r_obs[s * ldo + j] = src[j];
        }
        const int64_t k = ts_src[s];
        if (k >= 0) {
            for (int j = 0; j < ac; ++j) r_aacm[s * lda + j] = aacm[k * lda + j];
            r_oidx[s] = ts_oidx[s]; r_nidx[s] = ts_nidx[s];
            r_rew[s] = 0.f; r_done[s] = 0; r_end[s] = 0;      // ReplayBufferAcM.add_timestep stores no reward / done (replay_buffer.py:299-300)
        }
    }
}

// end[t * E + e] (float flags, step-major) -> out[e * T + t] (bytes, environment-major), with the batch cut at t = T - 1 counted as an end
__global__ void end_flags_env_major_kernel(const float* __restrict__ end, int E, int T, uint8_t* __restrict__ out) {
    __shared__ uint8_t tile[32][33];
    const int e0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, e = e0 + threadIdx.x;
        tile[i][threadIdx.x] = (t < T && e < E && (end[(size_t)t * E + e] != 0.f || t == T - 1)) ? 1 : 0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int e = e0 + i, t = t0 + threadIdx.x;
        if (e < E && t < T) out[(size_t)e * T + t] = tile[threadIdx.x][i];
    }
}

}  // namespace spp

namespace spp {

cudaError_t launch_end_flags_env_major(const float* end, int E, int T, uint8_t* out, cudaStream_t s) {
    dim3 grid((E + 31) / 32, (T + 31) / 32), block(32, 8);
    end_flags_env_major_kernel<<<grid, block, 0, s>>>(end, E, T, out);
    return cudaGetLastError();
}

cudaError_t launch_ring_add_store(float* r_obs, int32_t* r_oidx, int32_t* r_nidx, float* r_aacm, float* r_rew, uint8_t* r_done, uint8_t* r_end,
                                  int64_t S, int ob, int ac, int ldo, int lda, const int64_t* obs_src, const int64_t* ts_src,
                                  const int32_t* ts_oidx, const int32_t* ts_nidx, const PpoStoreView& st, cudaStream_t s) {
    const int blocks = (int)((S + 255) / 256);
    ring_add_store_kernel<<<blocks < 1 ? 1 : (blocks > 1184 ? 1184 : blocks), 256, 0, s>>>(r_obs, r_oidx, r_nidx, r_aacm, r_rew, r_done, r_end, S, ob, ac, ldo,
                                                                                   lda, obs_src, ts_src, ts_oidx, ts_nidx, st.raw_obs, st.raw_next, st.aacm);
    return cudaGetLastError();
}

cudaError_t launch_ring_gather(const RingView& R, int agent, const int64_t* d_idx, int n, const GatherOut& o, cudaStream_t s) {
    const int blocks = (n * 32 + 255) / 256;
    ring_gather_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, s>>>(R, agent, d_idx, n, o);
    return cudaGetLastError();
}

cudaError_t launch_ring_gather_bench(const RingView& R, int P, int nb, int B, const int64_t* d_len, uint64_t seed,
                                     float* out_obs, float* out_nobs, float* out_aacm, float* out_rew, uint8_t* out_done,
                                     int grid, cudaStream_t s) {
    ring_gather_bench_kernel<<<grid, 256, 0, s>>>(R, P, nb, B, d_len, seed, out_obs, out_nobs, out_aacm, out_rew, out_done);
    return cudaGetLastError();
}

cudaError_t launch_ring_fill(const RingView& R, float* obs, int32_t* oidx, int32_t* nidx, float* act, float* rew,
                             uint8_t* done, uint8_t* end, float* aacm, int P, int64_t n, int T, uint64_t seed,
                             const float* norm, int norm_stride, int grid, cudaStream_t s) {
    ring_fill_kernel<<<grid, 256, 0, s>>>(R, obs, oidx, nidx, act, rew, done, end, aacm, P, n, T, seed, norm, norm_stride);
    return cudaGetLastError();
}

}  // namespace spp
