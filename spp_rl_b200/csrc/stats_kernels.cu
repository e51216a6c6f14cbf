// Replay-ring observation statistics on the device: MetaReplayBuffer.update_obs_mean_std
// (rltoolkit/buffer/replay_buffer.py:83-96) = whole-buffer mean, population std (numpy, ddof 0) and the 1st / 99th percentiles
// (numpy "linear" interpolation between two order statistics) of `self.obs = _obs[_obs_idx[:current_len]]`, for every agent of a
// population at once.  The reference does this on the host with numpy every iteration (O(S log S) per agent and column).
//
// All passes are HBM-bound streams over the ring:  2 for the moments (sum, then sum of squared deviations from the mean, fp64,
// fixed summation order) and 4 for an exact 8-bit-per-pass radix select of four order statistics per column (the floor / ceil
// ranks of both percentiles), so every result is exactly an element of the buffer and the percentile interpolation is done
// in fp64 on the host with numpy's own formula -> bit-exact against np.percentile.
//
// Thread mapping (rows wider than 16 floats; narrower rows use the *_rows_kernel variants below): a block owns a contiguous range of rows; thread t < (256 / ob) * ob has a FIXED column j = t % ob and walks
// rows t / ob, t / ob + 256 / ob, ... of the range, so consecutive threads read consecutive floats of (mostly consecutive)
// ring rows and every thread keeps its accumulator / histogram column in registers / a private shared-memory slice.
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "ring_kernels.h"

namespace spp {

constexpr int kStatTargets = 4;      // order statistics per column: lo/hi rank of the 1st and of the 99th percentile
constexpr int kStatColsPerBlock = 12;
#ifndef SPP_STAT_ROWS
#define SPP_STAT_ROWS 8
#endif
constexpr int kStatRowsInFlight = SPP_STAT_ROWS;      // rows per thread and batch

__device__ __forceinline__ uint32_t sortable_key(float x) {      // monotone map float -> uint32
    const uint32_t b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ inline float key_to_float(uint32_t k) {
    const uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// pass 0: partial[a][blk][j] = sum of column j over the block's rows;  pass 1: sum of (x - mean[a][j])^2
__global__ void __launch_bounds__(256) stats_moment_kernel(RingView R, const int64_t* __restrict__ len, int pass,
                                                           const double* __restrict__ mean, double* __restrict__ partial) {
    const int a = blockIdx.y, nb = gridDim.x, ob = R.ob;
    const int64_t n = len[a];
    const int rows_per_iter = 256 / ob, active = rows_per_iter * ob;
    const int64_t per = (n + nb - 1) / nb, r0 = (int64_t)blockIdx.x * per, r1 = min(n, r0 + per);
    const size_t base = (size_t)a * R.S;
    __shared__ double red[256];
    double acc = 0.0;
    const int j = threadIdx.x % ob;
    if (threadIdx.x < active) {
        const double mu = pass ? mean[(size_t)a * ob + j] : 0.0;
        // 4 rows in flight per thread: the index load and the value load of a row are dependent, so without this every element
        // costs two exposed HBM latencies
        // software-pipelined: the indices of batch k + 1 are loaded while the values of batch k are in flight, so only ONE HBM
        // latency per batch is exposed (with the index load in front of its own value load the passes ran at 2.3 TB/s)
        constexpr int U = kStatRowsInFlight;
        int32_t on[U];
        int64_t i = r0 + threadIdx.x / ob;
#pragma unroll
        for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)u * rows_per_iter; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
        for (; i < r1; i += U * rows_per_iter) {
            int32_t oi[U]; float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { oi[u] = on[u]; v[u] = oi[u] >= 0 ? R.obs[(base + oi[u]) * R.ldo + j] : 0.f; }
#pragma unroll
            for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)(U + u) * rows_per_iter; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (oi[u] < 0) continue;
                const double x = (double)v[u];
                if (pass) { const double d = fabs(x - mu); acc += d * d; } else acc += x;
            }
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < ob) {      // fixed order: deterministic
        double s = 0.0;
        for (int t = threadIdx.x; t < active; t += ob) s += red[t];
        partial[((size_t)a * nb + blockIdx.x) * ob + threadIdx.x] = s;
    }
}

// Moment pass with a ROW per thread (observation rows of up to 16 floats: Pendulum, Hopper): float4 loads of the whole row, 4 rows in
// flight per thread (192 B instead of 16 B of loads outstanding per thread), one fp64 accumulator per column; the 256 per-thread
// sums of a column are combined by a fixed shuffle tree and the 8 warp partials in warp order (deterministic).
template <int NV>
__global__ void __launch_bounds__(256, 2) stats_moment_rows_kernel(RingView R, const int64_t* __restrict__ len, int pass,
                                                                const double* __restrict__ mean, double* __restrict__ partial) {
    const int a = blockIdx.y, nb = gridDim.x, ob = R.ob;
    const int64_t n = len[a];
    const int64_t per = (n + nb - 1) / nb, r0 = (int64_t)blockIdx.x * per, r1 = min(n, r0 + per);
    const size_t base = (size_t)a * R.S;
    constexpr int C = 4 * NV, U = 4;
    double acc[C], mu[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { acc[c] = 0.0; mu[c] = (pass && c < ob) ? mean[(size_t)a * ob + c] : 0.0; }
    int32_t on[U];
    int64_t i = r0 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)u * 256; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
    for (; i < r1; i += U * 256) {
        int32_t oi[U]; float4 v[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            oi[u] = on[u];
            const float4* row = reinterpret_cast<const float4*>(R.obs + (base + (oi[u] >= 0 ? oi[u] : 0)) * R.ldo);
#pragma unroll
            for (int q = 0; q < NV; ++q) v[u][q] = oi[u] >= 0 ? row[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)(U + u) * 256; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (oi[u] < 0) continue;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const float x4[4] = {v[u][q].x, v[u][q].y, v[u][q].z, v[u][q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double x = (double)x4[e];
                    if (pass) { const double d = fabs(x - mu[4 * q + e]); acc[4 * q + e] += d * d; } else acc[4 * q + e] += x;
                }
            }
        }
    }
    __shared__ double red[C][kWarps];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        double t = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane_id() == 0) red[c][warp_id()] = t;
    }
    __syncthreads();
    if (threadIdx.x < ob) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[threadIdx.x][w];
        partial[((size_t)a * nb + blockIdx.x) * ob + threadIdx.x] = t;
    }
}

static void launch_moment(const RingView& R, dim3 grid, const int64_t* d_len, int pass, const double* d_mean, double* d_partial, cudaStream_t s) {
    if (R.ldo == 4 && R.ob <= 4) stats_moment_rows_kernel<1><<<grid, 256, 0, s>>>(R, d_len, pass, d_mean, d_partial);
    else if (R.ldo == 8) stats_moment_rows_kernel<2><<<grid, 256, 0, s>>>(R, d_len, pass, d_mean, d_partial);
    else if (R.ldo == 12) stats_moment_rows_kernel<3><<<grid, 256, 0, s>>>(R, d_len, pass, d_mean, d_partial);
    else if (R.ldo == 16) stats_moment_rows_kernel<4><<<grid, 256, 0, s>>>(R, d_len, pass, d_mean, d_partial);
    else stats_moment_kernel<<<grid, 256, 0, s>>>(R, d_len, pass, d_mean, d_partial);      // wider rows: one column per thread
}

// out[a][j] = sum over blocks (fixed order) / n;  pass 1 additionally takes the square root
__global__ void stats_moment_finish_kernel(const double* __restrict__ partial, const int64_t* __restrict__ len, int P, int nb, int ob,
                                           int pass, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * ob) return;
    const int a = t / ob, j = t % ob;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[((size_t)a * nb + b) * ob + j];
    const double m = s / (double)len[a];
    out[t] = pass ? sqrt(m) : m;
}

// One radix pass (byte `pass` from the top): for every column and target, histogram the current byte of the keys whose higher
// bytes equal the target's prefix.  state[a][j][T] = {prefix (key bits found so far), remaining rank}; hist[a][j][T][256].
__global__ void __launch_bounds__(256) stats_select_hist_kernel(RingView R, const int64_t* __restrict__ len, int pass,
                                                                const unsigned long long* __restrict__ state, unsigned int* __restrict__ hist) {
    const int a = blockIdx.y, nb = gridDim.x, ob = R.ob;
    const int j0 = blockIdx.z * kStatColsPerBlock, cols = min(kStatColsPerBlock, ob - j0);
    const int64_t n = len[a];
    const int rows_per_iter = 256 / cols, active = rows_per_iter * cols;
    const int64_t per = (n + nb - 1) / nb, r0 = (int64_t)blockIdx.x * per, r1 = min(n, r0 + per);
    const size_t base = (size_t)a * R.S;
    extern __shared__ unsigned int sh[];      // [cols][kStatTargets][256]
    for (int i = threadIdx.x; i < cols * kStatTargets * 256; i += 256) sh[i] = 0;
    __syncthreads();
    if (threadIdx.x < active) {
        const int jj = threadIdx.x % cols, j = j0 + jj;
        const int shift = 24 - 8 * pass;
        uint32_t prefix[kStatTargets];
#pragma unroll
        for (int T = 0; T < kStatTargets; ++T) prefix[T] = (uint32_t)(state[(((size_t)a * ob + j) * kStatTargets + T) * 2]);
        unsigned int* h = sh + (size_t)jj * kStatTargets * 256;
        // targets that still share a prefix share a histogram (they differ only in the rank they look for)
        bool own[kStatTargets];
#pragma unroll
        for (int T = 0; T < kStatTargets; ++T) {
            own[T] = true;
#pragma unroll
            for (int U = 0; U < T; ++U) own[T] = own[T] && (prefix[U] != prefix[T]);
        }
        constexpr int U = kStatRowsInFlight;      // same software pipeline as the moment passes
        int32_t on[U];
        int64_t i = r0 + threadIdx.x / cols;
#pragma unroll
        for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)u * rows_per_iter; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
        for (; i < r1; i += U * rows_per_iter) {
            int32_t oi[U]; float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { oi[u] = on[u]; v[u] = oi[u] >= 0 ? R.obs[(base + oi[u]) * R.ldo + j] : 0.f; }
#pragma unroll
            for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)(U + u) * rows_per_iter; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (oi[u] < 0) continue;
                const uint32_t k = sortable_key(v[u]);
                const uint32_t byte = (k >> shift) & 255u;
                const uint32_t hi = pass ? (k >> (shift + 8)) : 0u;
#pragma unroll
                for (int T = 0; T < kStatTargets; ++T)
                    if (own[T] && (pass == 0 || hi == prefix[T])) atomicAdd(h + T * 256 + byte, 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cols * kStatTargets * 256; i += 256) {
        const unsigned int v = sh[i];
        if (v) atomicAdd(hist + ((size_t)a * ob + j0) * kStatTargets * 256 + i, v);
    }
}

// Radix pass with a ROW per thread (rows of up to 16 floats), same loads as stats_moment_rows_kernel.  All 32 lanes of a warp are in
// the same column at the same time, so hot bins serialise in the shared-memory atomic unit (pass 0: a handful of exponent bins);
// that costs less than the one-float-per-thread loads did (1.5-1.7 TB/s).  Later passes filter on the prefix first: few atomics.
template <int NV>
__global__ void __launch_bounds__(256, 2) stats_select_hist_rows_kernel(RingView R, const int64_t* __restrict__ len, int pass,
                                                                        const unsigned long long* __restrict__ state,
                                                                        unsigned int* __restrict__ hist) {
    const int a = blockIdx.y, nb = gridDim.x, ob = R.ob;
    const int64_t n = len[a];
    const int64_t per = (n + nb - 1) / nb, r0 = (int64_t)blockIdx.x * per, r1 = min(n, r0 + per);
    const size_t base = (size_t)a * R.S;
    constexpr int C = 4 * NV, U = 4;
    extern __shared__ unsigned int sh[];      // [ob][kStatTargets][256]
    for (int i = threadIdx.x; i < ob * kStatTargets * 256; i += 256) sh[i] = 0;
    const int shift = 24 - 8 * pass;
    uint32_t prefix[C][kStatTargets];         // registers: every thread sees every column
    unsigned long long ownbits = 0;           // bit 4 c + T: target T of column c owns a histogram (no earlier target shares its prefix)
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int T = 0; T < kStatTargets; ++T) prefix[c][T] = c < ob ? (uint32_t)(state[(((size_t)a * ob + c) * kStatTargets + T) * 2]) : 0u;
#pragma unroll
        for (int T = 0; T < kStatTargets; ++T) {
            bool own = c < ob;
#pragma unroll
            for (int V = 0; V < T; ++V) own = own && (prefix[c][V] != prefix[c][T]);
            if (own) ownbits |= 1ull << (4 * c + T);
        }
    }
    __syncthreads();
    int32_t on[U];
    int64_t i = r0 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)u * 256; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
    for (; i < r1; i += U * 256) {
        int32_t oi[U]; float4 v[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            oi[u] = on[u];
            const float4* row = reinterpret_cast<const float4*>(R.obs + (base + (oi[u] >= 0 ? oi[u] : 0)) * R.ldo);
#pragma unroll
            for (int q = 0; q < NV; ++q) v[u][q] = oi[u] >= 0 ? row[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int64_t ii = i + (int64_t)(U + u) * 256; on[u] = ii < r1 ? R.oidx[base + ii] : -1; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (oi[u] < 0) continue;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const float x4[4] = {v[u][q].x, v[u][q].y, v[u][q].z, v[u][q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * q + e;
                    if (c >= ob) continue;
                    const uint32_t k = sortable_key(x4[e]);
                    const uint32_t byte = (k >> shift) & 255u;
                    const uint32_t hi = pass ? (k >> (shift + 8)) : 0u;
#pragma unroll
                    for (int T = 0; T < kStatTargets; ++T)
                        if (((ownbits >> (4 * c + T)) & 1ull) && (pass == 0 || hi == prefix[c][T])) atomicAdd(sh + (c * kStatTargets + T) * 256 + byte, 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < ob * kStatTargets * 256; k += 256) {
        const unsigned int v = sh[k];
        if (v) atomicAdd(hist + (size_t)a * ob * kStatTargets * 256 + k, v);
    }
}

template <int NV>
static cudaError_t launch_hist_rows(const RingView& R, dim3 grid, const int64_t* d_len, int pass, const unsigned long long* d_state,
                                    unsigned int* d_hist, cudaStream_t s) {
    const size_t sh = (size_t)R.ob * kStatTargets * 256 * sizeof(unsigned int);
    cudaError_t e = cudaFuncSetAttribute(stats_select_hist_rows_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
    stats_select_hist_rows_kernel<NV><<<grid, 256, sh, s>>>(R, d_len, pass, d_state, d_hist);
    return cudaGetLastError();
}

// one warp per (agent, column): walk the 256 bins of every target, append the byte that contains its rank, clear the histogram
__global__ void stats_select_step_kernel(int P, int ob, int pass, unsigned long long* __restrict__ state, unsigned int* __restrict__ hist) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= P * ob) return;
    unsigned long long* st = state + (size_t)w * kStatTargets * 2;
    unsigned int* h = hist + (size_t)w * kStatTargets * 256;
    uint32_t prefix[kStatTargets];
    for (int T = 0; T < kStatTargets; ++T) prefix[T] = (uint32_t)st[2 * T];
    __syncwarp();
    for (int T = 0; T < kStatTargets; ++T) {
        int src = T;      // a target that shared its prefix with an earlier one reads that one's histogram
        for (int U = T - 1; U >= 0; --U) if (prefix[U] == prefix[T]) src = U;
        if (lane == 0) {
            unsigned long long rank = st[2 * T + 1], cum = 0;
            int b = 0;
            for (; b < 255; ++b) {
                const unsigned int c = h[src * 256 + b];
                if (cum + c > rank) break;
                cum += c;
            }
            st[2 * T] = ((unsigned long long)prefix[T] << 8) | (unsigned long long)b;
            st[2 * T + 1] = rank - cum;
        }
    }
    __syncwarp();
    for (int i = lane; i < kStatTargets * 256; i += 32) h[i] = 0;
}

float ring_stats_key_to_float(uint32_t key) { return key_to_float(key); }

cudaError_t launch_ring_obs_stats(const RingView& R, int P, const int64_t* d_len, const int64_t* h_len, int nb, double* d_partial,
                                  double* d_moments /*[2][P][ob]*/, unsigned long long* d_state, unsigned int* d_hist,
                                  unsigned long long* h_state /*pinned or pageable staging*/, cudaStream_t s) {
    const int ob = R.ob;
    double* d_mean = d_moments; double* d_std = d_moments + (size_t)P * ob;
    if (ob > 256) return cudaErrorInvalidValue;
    dim3 grid(nb, P);
    launch_moment(R, grid, d_len, 0, nullptr, d_partial, s);
    stats_moment_finish_kernel<<<(P * ob + 255) / 256, 256, 0, s>>>(d_partial, d_len, P, nb, ob, 0, d_mean);
    launch_moment(R, grid, d_len, 1, d_mean, d_partial, s);
    stats_moment_finish_kernel<<<(P * ob + 255) / 256, 256, 0, s>>>(d_partial, d_len, P, nb, ob, 1, d_std);
    // ranks as numpy computes them for method "linear": virtual index (n - 1) * q in fp64, lo = floor, hi = min(lo + 1, n - 1)
    for (int a = 0; a < P; ++a) {
        const int64_t n = h_len[a];
        const double q[2] = {0.01, 0.99};      // np.true_divide(1, 100), np.true_divide(99, 100)
        for (int j = 0; j < ob; ++j)
            for (int T = 0; T < kStatTargets; ++T) {
                const double v = (double)(n - 1) * q[T / 2];
                int64_t lo = (int64_t)floor(v);
                int64_t r = (T & 1) ? (lo + 1 < n ? lo + 1 : n - 1) : lo;
                if (r < 0) r = 0;
                h_state[(((size_t)a * ob + j) * kStatTargets + T) * 2] = 0ull;
                h_state[(((size_t)a * ob + j) * kStatTargets + T) * 2 + 1] = (unsigned long long)r;
            }
    }
    cudaError_t e = cudaMemcpyAsync(d_state, h_state, (size_t)P * ob * kStatTargets * 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_hist, 0, (size_t)P * ob * kStatTargets * 256 * sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    const int groups = (ob + kStatColsPerBlock - 1) / kStatColsPerBlock;
    const size_t sh = (size_t)kStatColsPerBlock * kStatTargets * 256 * sizeof(unsigned int);      // 48 KB
    e = cudaFuncSetAttribute(stats_select_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
    for (int pass = 0; pass < 4; ++pass) {
        if (R.ldo == 4 && ob <= 4) e = launch_hist_rows<1>(R, grid, d_len, pass, d_state, d_hist, s);
        else if (R.ldo == 8) e = launch_hist_rows<2>(R, grid, d_len, pass, d_state, d_hist, s);
        else if (R.ldo == 12) e = launch_hist_rows<3>(R, grid, d_len, pass, d_state, d_hist, s);
        else if (R.ldo == 16) e = launch_hist_rows<4>(R, grid, d_len, pass, d_state, d_hist, s);
        else stats_select_hist_kernel<<<dim3(nb, P, groups), 256, sh, s>>>(R, d_len, pass, d_state, d_hist);      // wider rows: one column per thread
        if (e != cudaSuccess) return e;
        stats_select_step_kernel<<<(P * ob * 32 + 255) / 256, 256, 0, s>>>(P, ob, pass, d_state, d_hist);
    }
    return cudaGetLastError();
}

}  // namespace spp
