// SPP-PPO critic fit on the 5th-generation tensor cores (A2C.update_critic, rltoolkit/algorithms/a2c/a2c.py:186-225; nets:
// rltoolkit/basic_model.py:64-77): one optimiser step's gradient of 0.5 mean((q - V(x))^2) over ALL rows of the rollout.
//
// Same contract as ppo_critic_grad_kernel (ppo_kernels.cu): every CTA owns a contiguous range of rows and leaves its partial
// gradient in its slot of `part` (+ the squared-error sum in `scal`); ppo_reduce_kernel adds the slots in a fixed order.
// What differs is where the work runs.  The FFMA version pushes 512-row blocks through five tile GEMMs whose operands travel
// through L2; here the weights are RESIDENT in shared memory, pre-split into the hi / lo tf32 planes of the three-pass scheme
// (umma.cuh), rows stream through in tiles of 128, and the three 64-wide contractions run on tcgen05:
//     fc2      D2[128 x 64]  = h1 . W2^T          A = h1  K-major (SWIZZLE_128B),       B = W2 K-major
//     dX fc2   D3[128 x 64]  = dz2 . W2           A = dz2 K-major,                      B = W2 MN-major (SWIZZLE_128B_BASE32B)
//     dW fc2   DW[64 x 64]  += dz2^T . h1         A = dz2 MN-major (M = 64),            B = h1 MN-major, K = rows
// fc1 (K = ob = 17) and its dW stay on FFMA from shared memory (4 % of the FLOPs each); tanh, the value head, the loss and
// the bias / fc3 gradients are row-wise register work between the products.
//
// Accuracy: an fp32 TMEM accumulator truncates, so every accumulator holds ONE 32-wide k-chunk (fc2, dX: the two
// halves of K = 64; dW: 32 rows) and the chunks are added in registers with round-to-nearest; cross terms are issued first.
// The dW accumulators are drained into per-thread running sums once per tile.  Measured tcgen05 facts this layout rests on
// (tools/umma_probe.py): M = 64 puts accumulator row 16 q + i on TMEM lane 32 q + i; a K-major operand cannot be read from an
// image in the BASE32B swizzle (device exception), so h1 and dz2 are written twice -- once per swizzle.
//
// Shared memory (209 KB, one CTA per SM): bufK 64 KB (h1, later dz2, K-major: 2 k-chunks x [hi | lo] x 128 rows x 128 B),
// W2 K-major 32 KB, W2 MN-major 32 KB, bufA / bufB 32 KB each (dz2 / h1 MN-major for HALF a tile: the dW product of a tile runs
// as two 64-row halves so that the images fit), x tile 10 KB, W1^T 5 KB, vectors.  TMEM: D2[2], D3[2], DW[4] x 64 columns.
#include <cstdlib>

#include "ppo_kernels.cuh"
#include "umma.cuh"
#include "update_kernel.cuh"      // NORM_* slots

namespace spp {

namespace {
using namespace umma;

constexpr int kTile = 128;
constexpr int kOffBufK = 0;
constexpr int kOffW2K = kOffBufK + 65536;
constexpr int kOffW2MN = kOffW2K + 32768;
constexpr int kOffBufA = kOffW2MN + 32768;
constexpr int kOffBufB = kOffBufA + 32768;
constexpr int kOffXs = kOffBufB + 32768;
constexpr int kXsLd = 20;                       // pad4(ob) for ob <= 20; pad columns are zero
constexpr int kXsBytes = kTile * kXsLd * 4;     // one x tile; two of them: the next tile lands (cp.async) while this one is processed
constexpr int kOffW1t = kOffXs + 2 * kXsBytes;
constexpr int kOffVec = kOffW1t + kXsLd * 64 * 4;
constexpr int kVecFloats = 64 * 3 + 4 * kTile + 32;     // b1, b2, w3, vpart[<= 4][128], reduction scratch
constexpr int kTcSmemBytes = kOffVec + kVecFloats * 4;
constexpr int kDz1Ld = 68;                      // raw dz1 tile (aliases bufA | bufB): conflict-free row writes and column reads
static_assert(kTile * kDz1Ld * 4 <= 65536, "dz1 tile aliases bufA | bufB");
static_assert(kTcSmemBytes + 1024 <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void issue3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc_first) {
    mma_tf32(d, al, bh, idesc, acc_first);
    mma_tf32(d, ah, bl, idesc, 1u);
}

// one 32-wide k-chunk of a three-pass product into accumulator d (fresh): cross terms of all four k-steps first, then hi x hi
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void chunk_mma(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t dah = A_MN ? mnmajor_desc(ah, ks) : kmajor_desc(ah, ks), dal = A_MN ? mnmajor_desc(al, ks) : kmajor_desc(al, ks);
        const uint64_t dbh = B_MN ? mnmajor_desc(bh, ks) : kmajor_desc(bh, ks), dbl = B_MN ? mnmajor_desc(bl, ks) : kmajor_desc(bl, ks);
        issue3(d, dah, dal, dbh, dbl, idesc, ks ? 1u : 0u);
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t dah = A_MN ? mnmajor_desc(ah, ks) : kmajor_desc(ah, ks);
        const uint64_t dbh = B_MN ? mnmajor_desc(bh, ks) : kmajor_desc(bh, ks);
        mma_tf32(d, dah, dbh, idesc, 1u);
    }
}

// this thread's CW values (columns c0 .. c0 + CW - 1 of row r) -> its row of the K-major tile pairs (hi, lo) of a [128 x 64] block:
// k-chunk c0 / 32 (32 KB apart: hi plane, then lo plane 16 KB further), sixteen-byte chunks (c0 % 32) / 4 ...
template <int CW>
__device__ __forceinline__ void store_kmajor_row(uint32_t buf, int r, int c0, const float (&v)[CW]) {
    const uint32_t hi_plane = buf + (c0 >> 5) * 32768, lo_plane = hi_plane + 16384;
    const int cb = (c0 & 31) >> 2;
#pragma unroll
    for (int c = 0; c < CW / 4; ++c) {
        float4 h, l;
        split_tf32(v[4 * c], h.x, l.x); split_tf32(v[4 * c + 1], h.y, l.y); split_tf32(v[4 * c + 2], h.z, l.z); split_tf32(v[4 * c + 3], h.w, l.w);
        const uint32_t off = kmajor_offset(r, cb + c);
        sts128(hi_plane + off, h); sts128(lo_plane + off, l);
    }
}
// ... and its k-row of an MN-major tile pair: k = row inside the 32-row chunk, MN group c0 / 32 (32 floats)
template <int CW>
__device__ __forceinline__ void store_mnmajor_row(uint32_t hi_plane, uint32_t lo_plane, int k, int c0, const float (&v)[CW]) {
    const int cb = 8 * (c0 >> 5) + ((c0 & 31) >> 2);
#pragma unroll
    for (int c = 0; c < CW / 4; ++c) {
        float4 h, l;
        split_tf32(v[4 * c], h.x, l.x); split_tf32(v[4 * c + 1], h.y, l.y); split_tf32(v[4 * c + 2], h.z, l.z); split_tf32(v[4 * c + 3], h.w, l.w);
        const uint32_t off = mnmajor_offset(k, cb + c);
        sts128(hi_plane + off, h); sts128(lo_plane + off, l);
    }
}

// the same values into BOTH images with one hi / lo split per element (dz2 of the first tile half goes to the K-major image for dX and
// to the MN-major image for dW in the same phase)
template <int CW>
__device__ __forceinline__ void store_both_rows(uint32_t kbuf, int r, uint32_t mn_hi, uint32_t mn_lo, int k, int c0, const float (&v)[CW]) {
    const uint32_t k_hi = kbuf + (c0 >> 5) * 32768, k_lo = k_hi + 16384;
    const int cbk = (c0 & 31) >> 2, cbm = 8 * (c0 >> 5) + ((c0 & 31) >> 2);
#pragma unroll
    for (int c = 0; c < CW / 4; ++c) {
        float4 h, l;
        split_tf32(v[4 * c], h.x, l.x); split_tf32(v[4 * c + 1], h.y, l.y); split_tf32(v[4 * c + 2], h.z, l.z); split_tf32(v[4 * c + 3], h.w, l.w);
        const uint32_t ok = kmajor_offset(r, cbk + c), om = mnmajor_offset(k, cbm + c);
        sts128(k_hi + ok, h); sts128(k_lo + ok, l);
        sts128(mn_hi + om, h); sts128(mn_lo + om, l);
    }
}

// (d0, d1) += a * (b0, b1): one packed FFMA2 (sm_100 fma.rn.f32x2; each half is an IEEE fma, so results equal two scalar fmaf)
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}

// v[j] (+)= TMEM[lane][col0 + j], j < CW (32 or 16)
template <bool ADD, int CW>
__device__ __forceinline__ void tmem_row(uint32_t taddr, float (&v)[CW]) {
    float t0[16];
    tmem_ld16_nowait(taddr, t0);
    if constexpr (CW == 32) {
        float t1[16];
        tmem_ld16_nowait(taddr + 16, t1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 + j] = ADD ? __fadd_rn(v[16 + j], t1[j]) : t1[j];
    } else {
        tmem_ld_wait();
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ADD ? __fadd_rn(v[j], t0[j]) : t0[j];
}
template <bool ADD>
__device__ __forceinline__ void tmem_row32(uint32_t taddr, float (&v)[32]) { tmem_row<ADD, 32>(taddr, v); }

// block-wide sum in a fixed order for NW warps (common.cuh's block_sum is written for the 8-warp CTAs)
template <int NW>
__device__ __forceinline__ float block_sum_n(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[warp_id()] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += red[w];
    return t;
}
}  // namespace

// NG column groups per row: a thread owns one row x CW = 64 / NG columns; 128 NG threads (8 or 16 warps).
template <int NG>
__global__ void __launch_bounds__(128 * NG, 1) ppo_critic_grad_tc_kernel(const __grid_constant__ PpoArgs a) {
    constexpr int kThr = 128 * NG, CW = 64 / NG;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t s0 = smem_u32(smem);
    float* xs_base = reinterpret_cast<float*>(smem + kOffXs);
    float* w1t = reinterpret_cast<float*>(smem + kOffW1t);
    float* vec = reinterpret_cast<float*>(smem + kOffVec);
    float* b1s = vec; float* b2s = vec + 64; float* w3s = vec + 128; float* vpart = vec + 192; float* red = vec + 192 + 4 * kTile;
    float* dz1s = reinterpret_cast<float*>(smem + kOffBufA);
    __shared__ uint64_t mbar3[3];      // fc2 | dW fc2 (both halves) | dX
    __shared__ uint32_t tmem_base_s;

    float* part = a.part + (size_t)blockIdx.x * a.part_stride;
    float* scal = a.scal + (size_t)blockIdx.x * PS_COUNT;
    for (int i = threadIdx.x; i < a.L.critic.size; i += kThr) part[i] = 0.f;
    if (threadIdx.x < PS_COUNT) scal[threadIdx.x] = 0.f;
    const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_cta;
    int64_t nrows64 = a.d.N - r0;
    if (nrows64 > a.rows_per_cta) nrows64 = a.rows_per_cta;
    if (nrows64 <= 0) return;
    const int nrows = (int)nrows64;

    const LayerDesc& l0 = a.L.critic.L[0]; const LayerDesc& l1 = a.L.critic.L[1]; const LayerDesc& l2 = a.L.critic.L[2];
    const int ob = a.L.ob, ldo = a.L.ldo;

    if (warp_id() == 0) tmem_alloc<512>(&tmem_base_s);
    if (threadIdx.x == 0) { mbar_init(mbar3, 1); mbar_init(mbar3 + 1, 1); mbar_init(mbar3 + 2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    // resident weights: W1^T (fp32, FFMA), b1, b2, w3; W2 split into hi / lo planes in both operand layouts
    for (int e = threadIdx.x; e < kXsLd * 64; e += kThr) {
        const int i = e / 64, c = e % 64;
        w1t[e] = (i < ob) ? a.critic[l0.off_w + c * l0.ld + i] : 0.f;
    }
    if (threadIdx.x < 64) {
        b1s[threadIdx.x] = a.critic[l0.off_b + threadIdx.x]; b2s[threadIdx.x] = a.critic[l1.off_b + threadIdx.x];
        w3s[threadIdx.x] = a.critic[l2.off_w + threadIdx.x];
    }
    for (int e = threadIdx.x; e < 64 * 16; e += kThr) {      // float4 (o, i4) of W2 [out][in]
        const int o = e >> 4, i4 = e & 15;
        const float4 w = *reinterpret_cast<const float4*>(a.critic + l1.off_w + o * l1.ld + 4 * i4);
        float4 h, l;
        split_tf32(w.x, h.x, l.x); split_tf32(w.y, h.y, l.y); split_tf32(w.z, h.z, l.z); split_tf32(w.w, h.w, l.w);
        {   // fc2: B[n = out][k = in], K-major; k-chunk = in / 32
            const uint32_t base = s0 + kOffW2K + (i4 >> 3) * 16384, off = kmajor_offset(o, i4 & 7);
            sts128(base + off, h); sts128(base + 8192 + off, l);
        }
        {   // dX: B[k = out][n = in], MN-major; k-chunk = out / 32
            const uint32_t base = s0 + kOffW2MN + (o >> 5) * 16384, off = mnmajor_offset(o & 31, i4);
            sts128(base + off, h); sts128(base + 8192 + off, l);
        }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const float b3 = a.critic[l2.off_b];
    const float inv_n = 1.0f / (float)a.d.Ntot;

    const int q = warp_id() & 3, hcol = warp_id() >> 2, c0 = CW * hcol;
    const int r = 32 * q + lane_id();                                    // this thread's row of the tile
    const uint32_t my_t = tmem + ((uint32_t)(32 * q) << 16) + c0;         // its TMEM lane quarter / column half
    const uint32_t idesc_fwd = make_idesc_tf32(128, 64, 0, 0), idesc_dx = make_idesc_tf32(128, 64, 0, 1), idesc_dw = make_idesc_tf32(64, 64, 1, 1);
    uint32_t ph_fc2 = 0, ph_dw = 0, ph_dx = 0;
    uint64_t* const m_fc2 = mbar3; uint64_t* const m_dw = mbar3 + 1; uint64_t* const m_dx = mbar3 + 2;

    constexpr int kCB = NG == 2 ? 5 : 3;      // float4 column blocks of dW fc1 per thread
    float gw2[CW], cs_b2[CW], cs_w3[CW], gw1[4 * kCB];      // running sums over the CTA's tiles
#pragma unroll
    for (int j = 0; j < CW; ++j) { gw2[j] = 0.f; cs_b2[j] = 0.f; cs_w3[j] = 0.f; }
#pragma unroll
    for (int m = 0; m < 4 * kCB; ++m) gw1[m] = 0.f;
    float sse = 0.f, sdv = 0.f, gb1 = 0.f;
    // dW1 mapping: out unit o1, rows 32 rq .. 32 rq + 31, float4 column blocks cb0 .. cb0 + ncb - 1 (8 warps: all five; 16 warps: 3 + 2)
    const int o1 = threadIdx.x & 63;
    const int cg = NG == 2 ? 0 : (threadIdx.x >> 6) & 1, rq = NG == 2 ? threadIdx.x >> 6 : threadIdx.x >> 7;
    const int cb0 = 3 * cg, ncb = NG == 2 ? 5 : (cg ? 2 : 3);

    // x tile t -> shared buffer t & 1 by 16-byte cp.async (rows beyond the range and pad chunks zero-filled)
    auto load_x_tile = [&](int t0, float* dst) {
        const int trows = min(kTile, nrows - t0);
        const float* X = a.d.x + (r0 + t0) * ldo;
        for (int e = threadIdx.x; e < kTile * (kXsLd / 4); e += kThr) {
            const int rr = e / (kXsLd / 4), c4 = e % (kXsLd / 4);
            const bool ok = rr < trows && 4 * c4 < ldo;
            cp_async16(dst + rr * kXsLd + 4 * c4, ok ? X + (size_t)rr * ldo + 4 * c4 : a.d.x, ok ? 16 : 0);
        }
        cp_async_commit();
    };
    // dW fc1 (+ d b1) on FFMA from the raw dz1 tile (bufA | bufB) and an x tile: out unit o1 x all input columns over this thread's
    // quarter of the rows.  Runs for tile t - 1 inside tile t's fc2 wait (and once after the loop).
    auto dw1_ffma = [&](const float* xs) {
#pragma unroll 4
        for (int rr = 32 * rq; rr < 32 * rq + 32; ++rr) {
            const float d = dz1s[rr * kDz1Ld + o1];
            if (cg == 0) gb1 = __fadd_rn(gb1, d);
            const float4* xr = reinterpret_cast<const float4*>(xs + rr * kXsLd) + cb0;      // warp-uniform address: broadcast
#pragma unroll
            for (int c = 0; c < kCB; ++c) {
                if (c >= ncb) continue;
                const float4 x4 = xr[c];
                fma2(gw1[4 * c], gw1[4 * c + 1], d, x4.x, x4.y);
                fma2(gw1[4 * c + 2], gw1[4 * c + 3], d, x4.z, x4.w);
            }
        }
    };
    load_x_tile(0, xs_base);
    int tbuf = 0;
    for (int t0 = 0; t0 < nrows; t0 += kTile, tbuf ^= 1) {
        const int trows = min(kTile, nrows - t0);
        float* xs = xs_base + tbuf * (kXsBytes / 4);
        cp_async_wait<0>();
        __syncthreads();
        const float q_row = (r < trows) ? __ldg(a.d.q + r0 + t0 + r) : 0.f;      // needed by the value head: in flight during fc1 / fc2
        // ---- fc1 (FFMA): h1 = tanh(x W1^T + b1), this thread's row, 32 columns
        float h1[CW];
        {
            float acc[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) acc[j] = 0.f;
            float xv[kXsLd];      // the row's 20 floats with conflict-free 16-byte loads; pad columns meet zero rows of W1^T
#pragma unroll
            for (int c = 0; c < kXsLd / 4; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(xs + r * kXsLd + 4 * c);
                xv[4 * c] = v.x; xv[4 * c + 1] = v.y; xv[4 * c + 2] = v.z; xv[4 * c + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < kXsLd; ++i) {
                const float4* wr = reinterpret_cast<const float4*>(w1t + i * 64 + c0);
#pragma unroll
                for (int c = 0; c < CW / 4; ++c) {
                    const float4 w = wr[c];
                    fma2(acc[4 * c], acc[4 * c + 1], xv[i], w.x, w.y);
                    fma2(acc[4 * c + 2], acc[4 * c + 3], xv[i], w.z, w.w);
                }
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) h1[j] = tanhf(__fadd_rn(acc[j], b1s[c0 + j]));
        }
        store_kmajor_row<CW>(s0 + kOffBufK, r, c0, h1);
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {      // fc2
            fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<false, false>(tmem + 64 * c, s0 + kOffBufK + c * 32768, s0 + kOffBufK + c * 32768 + 16384,
                                        s0 + kOffW2K + c * 16384, s0 + kOffW2K + c * 16384 + 8192, idesc_fwd);
            commit(m_fc2);
        }
        if (t0 > 0) dw1_ffma(xs_base + (tbuf ^ 1) * (kXsBytes / 4));      // fc1 weight gradient of the PREVIOUS tile while fc2 runs
        mbar_wait(m_fc2, ph_fc2); ph_fc2 ^= 1;
        fence_after_sync();
        // ---- value head, loss, dz2
        float dz2[CW];
        float dv;
        {
            float h2[CW];
            tmem_row<false, CW>(my_t, h2);
            tmem_row<true, CW>(my_t + 64, h2);
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < CW; ++j) { h2[j] = tanhf(__fadd_rn(h2[j], b2s[c0 + j])); dot = fmaf(h2[j], w3s[c0 + j], dot); }
            vpart[hcol * kTile + r] = dot;
            __syncthreads();      // also: every thread is past the previous tile's dW1 loop -- its x buffer and bufA | bufB are free
            if (t0 + kTile < nrows) load_x_tile(t0 + kTile, xs_base + (tbuf ^ 1) * (kXsBytes / 4));      // lands during the rest of this tile
            float v = vpart[r];
#pragma unroll
            for (int g = 1; g < NG; ++g) v = __fadd_rn(v, vpart[g * kTile + r]);
            v = __fadd_rn(v, b3);
            const bool valid = r < trows;
            const float diff = valid ? __fsub_rn(q_row, v) : 0.f;
            dv = -__fmul_rn(diff, inv_n);
            if (hcol == 0) { sse = fmaf(diff, diff, sse); sdv += dv; }
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                dz2[j] = __fmul_rn(__fmul_rn(dv, w3s[c0 + j]), __fsub_rn(1.f, __fmul_rn(h2[j], h2[j])));
                cs_b2[j] += dz2[j];
                cs_w3[j] = fmaf(dv, h2[j], cs_w3[j]);
            }
        }
        // dz2 -> K-major (over h1's image: the fc2 products have retired); first half of the tile -> MN-major images of dz2 and h1
        if (q < 2) {
            const uint32_t pa = s0 + kOffBufA + (r >> 5) * 16384, pb = s0 + kOffBufB + (r >> 5) * 16384;
            store_both_rows<CW>(s0 + kOffBufK, r, pa, pa + 8192, r & 31, c0, dz2);
            store_mnmajor_row<CW>(pb, pb + 8192, r & 31, c0, h1);
        } else {
            store_kmajor_row<CW>(s0 + kOffBufK, r, c0, dz2);
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {      // dW fc2 of rows 0..63 first (its images are needed back soonest), then dX through fc2 (all 128 rows)
            fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<true, true>(tmem + 256 + 64 * c, s0 + kOffBufA + c * 16384, s0 + kOffBufA + c * 16384 + 8192,
                                      s0 + kOffBufB + c * 16384, s0 + kOffBufB + c * 16384 + 8192, idesc_dw);
            commit(m_dw);
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<false, true>(tmem + 128 + 64 * c, s0 + kOffBufK + c * 32768, s0 + kOffBufK + c * 32768 + 16384,
                                       s0 + kOffW2MN + c * 16384, s0 + kOffW2MN + c * 16384 + 8192, idesc_dx);
            commit(m_dx);
        }
        mbar_wait(m_dw, ph_dw); ph_dw ^= 1;
        fence_after_sync();
        if (q >= 2) {                // second half of the tile -> the MN-major images
            const int lr = r - 64;
            const uint32_t pa = s0 + kOffBufA + (lr >> 5) * 16384, pb = s0 + kOffBufB + (lr >> 5) * 16384;
            store_mnmajor_row<CW>(pa, pa + 8192, lr & 31, c0, dz2);
            store_mnmajor_row<CW>(pb, pb + 8192, lr & 31, c0, h1);
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {      // dW fc2 of rows 64..127
            fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<true, true>(tmem + 384 + 64 * c, s0 + kOffBufA + c * 16384, s0 + kOffBufA + c * 16384 + 8192,
                                      s0 + kOffBufB + c * 16384, s0 + kOffBufB + c * 16384 + 8192, idesc_dw);
            commit(m_dw);
        }
        // ---- dz1 = (dz2 W2) (1 - h1^2) while the tensor core finishes the second half
        mbar_wait(m_dx, ph_dx); ph_dx ^= 1;
        fence_after_sync();
        float dz1[CW];
        tmem_row<false, CW>(my_t + 128, dz1);
        tmem_row<true, CW>(my_t + 192, dz1);
#pragma unroll
        for (int j = 0; j < CW; ++j) dz1[j] = __fmul_rn(dz1[j], __fsub_rn(1.f, __fmul_rn(h1[j], h1[j])));
        mbar_wait(m_dw, ph_dw); ph_dw ^= 1;
        fence_after_sync();
        // raw dz1 tile over bufA | bufB (their products have retired) for the fc1 weight gradient
#pragma unroll
        for (int c = 0; c < CW / 4; ++c)
            *reinterpret_cast<float4*>(dz1s + r * kDz1Ld + c0 + 4 * c) = make_float4(dz1[4 * c], dz1[4 * c + 1], dz1[4 * c + 2], dz1[4 * c + 3]);
        // dW fc2: the four 32-row accumulators of this tile -> running sums (accumulator row 16 q + i sits on lane 32 q + i)
        {
            float t[CW];
            tmem_row<false, CW>(my_t + 256, t);
            tmem_row<true, CW>(my_t + 320, t);
            tmem_row<true, CW>(my_t + 384, t);
            tmem_row<true, CW>(my_t + 448, t);
#pragma unroll
            for (int j = 0; j < CW; ++j) gw2[j] = __fadd_rn(gw2[j], t[j]);
        }
        fence_before_sync();      // (the next tile's barriers order these TMEM reads before its MMAs)
    }
    __syncthreads();
    dw1_ffma(xs_base + (tbuf ^ 1) * (kXsBytes / 4));      // the last tile (tbuf was flipped by the loop increment)

    // ---- the CTA's partial gradient
    if (lane_id() < 16) {
        const int o = 16 * q + lane_id();
#pragma unroll
        for (int c = 0; c < CW / 4; ++c)
            *reinterpret_cast<float4*>(part + l1.off_w + o * l1.ld + c0 + 4 * c) = make_float4(gw2[4 * c], gw2[4 * c + 1], gw2[4 * c + 2], gw2[4 * c + 3]);
    }
    {   // the four row quarters of dW fc1, added in quarter order
        float* w1red = reinterpret_cast<float*>(smem + kOffBufA);      // [4][64][21]: 20 weight columns + the bias
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4 * kCB; ++m)
            if (m < 4 * ncb) w1red[(rq * 64 + o1) * 21 + 4 * cb0 + m] = gw1[m];
        if (cg == 0) w1red[(rq * 64 + o1) * 21 + kXsLd] = gb1;
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * 21; e += kThr) {
            const int o = e / 21, i = e % 21;
            const float sum = __fadd_rn(__fadd_rn(__fadd_rn(w1red[(0 * 64 + o) * 21 + i], w1red[(1 * 64 + o) * 21 + i]), w1red[(2 * 64 + o) * 21 + i]),
                                        w1red[(3 * 64 + o) * 21 + i]);
            if (i < ob) part[l0.off_w + o * l0.ld + i] = sum;
            else if (i == kXsLd) part[l0.off_b + o] = sum;
        }
    }
    // column sums over the 128 row-threads of each column half, in row order (deterministic): d b2 and d w3
    float* colred = reinterpret_cast<float*>(smem + kOffBufK);      // [NG][128][CW + 1]
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CW; ++j) colred[(hcol * kTile + r) * (CW + 1) + j] = pass == 0 ? cs_b2[j] : cs_w3[j];
        __syncthreads();
        if (threadIdx.x < 64) {
            const int hh = threadIdx.x / CW, j = threadIdx.x % CW;
            float s = 0.f;
            for (int rr = 0; rr < kTile; ++rr) s += colred[(hh * kTile + rr) * (CW + 1) + j];
            part[(pass == 0 ? l1.off_b : l2.off_w) + threadIdx.x] = s;
        }
    }
    const float t1 = block_sum_n<kThr / 32>(sse, red);
    __syncthreads();
    const float t2 = block_sum_n<kThr / 32>(sdv, red);
    if (threadIdx.x == 0) { scal[PS_LOSS] = t1; part[l2.off_b] = t2; }
    fence_before_sync();
    __syncthreads();
    if (warp_id() == 0) tmem_dealloc<512>(tmem);
}

cudaError_t launch_ppo_critic_grad_tc(const PpoArgs& a, int grid, cudaStream_t s) {
    if (!ppo_critic_tc_supported(a.L.ob, a.L.ldo)) return cudaErrorInvalidValue;      // wider observations take the FFMA kernel
    // 8 warps (a thread = one row x 32 columns, 255 registers; default) or 16 (one row x 16 columns, 128 registers: SPP_PPO_CRITIC_WARPS=16).
    // Measured equal (4.80 vs 4.88 ms per step at 8.39 M rows): twice the warps with half the instruction-level parallelism each.
    static const int warps = [] { const char* e = getenv("SPP_PPO_CRITIC_WARPS"); return (e && atoi(e) == 16) ? 16 : 8; }();
    cudaError_t e;
    if (warps == 16) {
        e = cudaFuncSetAttribute(ppo_critic_grad_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes + 1024);
        if (e != cudaSuccess) return e;
        ppo_critic_grad_tc_kernel<4><<<grid, 512, kTcSmemBytes + 1024, s>>>(a);
    } else {
        e = cudaFuncSetAttribute(ppo_critic_grad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes + 1024);
        if (e != cudaSuccess) return e;
        ppo_critic_grad_tc_kernel<2><<<grid, 256, kTcSmemBytes + 1024, s>>>(a);
    }
    return cudaGetLastError();
}

bool ppo_critic_tc_supported(int ob, int ldo) { return ob <= kXsLd && ldo <= kXsLd; }

// ---- forward only: V(x) -> v, V(xn) -> nv, q = r + gamma (1 - done) nv (ppo_critic_values_kernel's contract, modes 0 / 1) with fc2 on
// tcgen05.  Shared memory: h1 K-major 64 KB + W2 K-major 32 KB + one x tile + W1^T + vectors = 111 KB, 128 TMEM columns: two CTAs
// per SM, so one CTA's row-wise work covers the other's tensor-core wait.
namespace {
constexpr int kValOffW2K = 65536;
constexpr int kValOffW1t = kValOffW2K + 32768;
constexpr int kValOffVec = kValOffW1t + kXsLd * 64 * 4;
constexpr int kValSmemBytes = kValOffVec + (64 * 3 + 2 * kTile) * 4;
static_assert(2 * (kValSmemBytes + 1024 + 1024) <= 227 * 1024, "two value CTAs per SM");
}  // namespace

__global__ void __launch_bounds__(kThreads, 2) ppo_critic_values_tc_kernel(const __grid_constant__ PpoArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t s0 = smem_u32(smem);
    float* w1t = reinterpret_cast<float*>(smem + kValOffW1t);
    float* vec = reinterpret_cast<float*>(smem + kValOffVec);
    float* b1s = vec; float* b2s = vec + 64; float* w3s = vec + 128; float* vpart = vec + 192;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_cta;
    int64_t nrows64 = a.d.N - r0;
    if (nrows64 > a.rows_per_cta) nrows64 = a.rows_per_cta;
    if (nrows64 <= 0) return;
    const int nrows = (int)nrows64;
    const LayerDesc& l0 = a.L.critic.L[0]; const LayerDesc& l1 = a.L.critic.L[1]; const LayerDesc& l2 = a.L.critic.L[2];
    const int ob = a.L.ob, ldo = a.L.ldo;
    if (warp_id() == 0) tmem_alloc<128>(&tmem_base_s);
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int e = threadIdx.x; e < kXsLd * 64; e += kThreads) {
        const int i = e / 64, c = e % 64;
        w1t[e] = (i < ob) ? a.critic[l0.off_w + c * l0.ld + i] : 0.f;
    }
    if (threadIdx.x < 64) {
        b1s[threadIdx.x] = a.critic[l0.off_b + threadIdx.x]; b2s[threadIdx.x] = a.critic[l1.off_b + threadIdx.x];
        w3s[threadIdx.x] = a.critic[l2.off_w + threadIdx.x];
    }
    for (int e = threadIdx.x; e < 64 * 16; e += kThreads) {
        const int o = e >> 4, i4 = e & 15;
        const float4 w = *reinterpret_cast<const float4*>(a.critic + l1.off_w + o * l1.ld + 4 * i4);
        float4 h, l;
        split_tf32(w.x, h.x, l.x); split_tf32(w.y, h.y, l.y); split_tf32(w.z, h.z, l.z); split_tf32(w.w, h.w, l.w);
        const uint32_t base = s0 + kValOffW2K + (i4 >> 3) * 16384, off = kmajor_offset(o, i4 & 7);
        sts128(base + off, h); sts128(base + 8192 + off, l);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const float b3 = a.critic[l2.off_b];
    const int q = warp_id() & 3, hcol = warp_id() >> 2, c0 = 32 * hcol;
    const int r = 32 * q + lane_id();
    const uint32_t my_t = tmem + ((uint32_t)(32 * q) << 16) + c0;
    const uint32_t idesc_fwd = make_idesc_tf32(128, 64, 0, 0);
    uint32_t phase = 0;
    for (int pass = (a.mode == 1 ? 1 : 0); pass < 2; ++pass) {
        const float* src = pass == 0 ? a.d.x : a.d.xn;
        float* dst = pass == 0 ? a.d.v : a.d.nv;
        for (int t0 = 0; t0 < nrows; t0 += kTile) {
            const int trows = min(kTile, nrows - t0);
            const float* X = src + (r0 + t0) * ldo;
            float h[32];
            {
                float acc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0.f;
                float xv[kXsLd];      // this thread's row straight from global memory (both column halves read it: the second hits L1)
#pragma unroll
                for (int c = 0; c < kXsLd / 4; ++c) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < trows && 4 * c < ldo) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)r * ldo + 4 * c));
                    xv[4 * c] = v.x; xv[4 * c + 1] = v.y; xv[4 * c + 2] = v.z; xv[4 * c + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < kXsLd; ++i) {
                    const float4* wr = reinterpret_cast<const float4*>(w1t + i * 64 + c0);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 w = wr[c];
                        fma2(acc[4 * c], acc[4 * c + 1], xv[i], w.x, w.y);
                        fma2(acc[4 * c + 2], acc[4 * c + 3], xv[i], w.z, w.w);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) h[j] = tanhf(__fadd_rn(acc[j], b1s[c0 + j]));
            }
            store_kmajor_row<32>(s0, r, c0, h);
            fence_proxy_async();
            fence_before_sync();
            __syncthreads();
            if (threadIdx.x == 0) {
                fence_after_sync();
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    chunk_mma<false, false>(tmem + 64 * c, s0 + c * 32768, s0 + c * 32768 + 16384, s0 + kValOffW2K + c * 16384,
                                            s0 + kValOffW2K + c * 16384 + 8192, idesc_fwd);
                commit(&mbar);
            }
            mbar_wait(&mbar, phase); phase ^= 1;
            fence_after_sync();
            tmem_row32<false>(my_t, h);
            tmem_row32<true>(my_t + 64, h);
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) dot = fmaf(tanhf(__fadd_rn(h[j], b2s[c0 + j])), w3s[c0 + j], dot);
            vpart[hcol * kTile + r] = dot;
            fence_before_sync();
            __syncthreads();
            if (hcol == 0 && r < trows) dst[r0 + t0 + r] = __fadd_rn(__fadd_rn(vpart[r], vpart[kTile + r]), b3);
            // (the next tile's barrier before its MMA separates these vpart reads from its writes)
        }
    }
    __syncthreads();      // nv of this CTA's rows is complete and visible to the CTA
    for (int i = threadIdx.x; i < nrows; i += kThreads) {
        const int64_t rr = r0 + i;
        a.d.q[rr] = __fadd_rn(a.d.rew[rr], __fmul_rn(__fmul_rn(a.h.gamma, __fsub_rn(1.f, a.d.done[rr])), a.d.nv[rr]));
    }
    fence_before_sync();
    __syncthreads();
    if (warp_id() == 0) tmem_dealloc<128>(tmem);
}

cudaError_t launch_ppo_critic_values_tc(const PpoArgs& a, int grid, cudaStream_t s) {
    if (!ppo_critic_tc_supported(a.L.ob, a.L.ldo)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(ppo_critic_values_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kValSmemBytes + 1024);
    if (e != cudaSuccess) return e;
    ppo_critic_values_tc_kernel<<<grid, kThreads, kValSmemBytes + 1024, s>>>(a);
    return cudaGetLastError();
}



// =====================================================================================================================================
// The clipped-ratio actor step (PPO_AcM.update_actor_acm / PPO.update_actor / the A2C forms; ppo_actor_grad_kernel's contract: one
// gathered minibatch in a.b, partial gradients per CTA) with the same tile machinery: fc2, dX(fc2) and dW(fc2) on tcgen05, W2 resident
// as pre-split planes; fc1, the 17-wide head (fc3, Gaussian log-prob, ratio / clip, its backward) and dW(fc1) / dW(fc3) on FFMA from
// shared memory.  Per tile, between the fc2 product and the dz2 images, bufA | bufB hold the head's scratch: the two column halves'
// partial fc3 pre-activations, d3 and the log-scale terms per row, and the raw h2 tile for dW(fc3).
namespace {
constexpr int kActOffW3 = kOffVec + kVecFloats * 4;                 // W3 [20][64] (rows >= ob zero)
constexpr int kActOffHv = kActOffW3 + kXsLd * 64 * 4;               // b3, lim, var, 2 var, log sd [5][20]
constexpr int kActSmemBytes = kActOffHv + 5 * kXsLd * 4;
static_assert(kActSmemBytes + 1024 <= 227 * 1024, "shared memory budget (actor): dynamic + the kernel's static 1 KB");
constexpr int kExOff = 0;                                           // floats inside bufA | bufB: exch [2][128][20]
constexpr int kD3Off = 2 * kTile * kXsLd;                           // d3s [128][20]
constexpr int kH2Off = 3 * kTile * kXsLd;                           // h2s [128][68]
static_assert((kH2Off + kTile * kDz1Ld) * 4 <= 65536, "head scratch aliases bufA | bufB");
}  // namespace

__global__ void __launch_bounds__(kThreads, 1) ppo_actor_grad_tc_kernel(const __grid_constant__ PpoArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw;
    const uint32_t s0 = smem_u32(smem);
    if (s0 & 1023u) __trap();
    float* xs_base = reinterpret_cast<float*>(smem + kOffXs);
    float* w1t = reinterpret_cast<float*>(smem + kOffW1t);
    float* vec = reinterpret_cast<float*>(smem + kOffVec);
    float* b1s = vec; float* b2s = vec + 64; float* red = vec + 192 + 4 * kTile;
    float* w3s = reinterpret_cast<float*>(smem + kActOffW3);
    float* hv = reinterpret_cast<float*>(smem + kActOffHv);
    float* b3v = hv; float* limv = hv + kXsLd; float* varv = hv + 2 * kXsLd; float* var2v = hv + 3 * kXsLd; float* logsdv = hv + 4 * kXsLd;
    float* scr = reinterpret_cast<float*>(smem + kOffBufA);          // head scratch / raw dz1 tile (bufA | bufB)
    float* dz1s = scr; float* exch = scr + kExOff; float* d3s = scr + kD3Off; float* h2s = scr + kH2Off;
    __shared__ uint64_t mbar3[3];
    __shared__ uint32_t tmem_base_s;

    float* part = a.part + (size_t)blockIdx.x * a.part_stride;
    float* scal = a.scal + (size_t)blockIdx.x * PS_COUNT;
    for (int i = threadIdx.x; i < a.L.actor.size; i += kThreads) part[i] = 0.f;
    if (threadIdx.x < PS_COUNT) scal[threadIdx.x] = 0.f;
    const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_cta;
    int64_t nrows64 = a.b.n - r0;
    if (nrows64 > a.rows_per_cta) nrows64 = a.rows_per_cta;
    if (nrows64 <= 0) return;
    const int nrows = (int)nrows64;
    const LayerDesc& l0 = a.L.actor.L[0]; const LayerDesc& l1 = a.L.actor.L[1]; const LayerDesc& l2 = a.L.actor.L[2]; const LayerDesc& l3 = a.L.actor.L[3];
    const int ob = a.L.ob, ldo = a.L.ldo;
    const float* lim = a.norm + NORM_LIM * ldo;

    if (warp_id() == 0) tmem_alloc<512>(&tmem_base_s);
    if (threadIdx.x == 0) { mbar_init(mbar3, 1); mbar_init(mbar3 + 1, 1); mbar_init(mbar3 + 2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int e = threadIdx.x; e < kXsLd * 64; e += kThreads) {
        const int i = e / 64, c = e % 64;
        w1t[e] = (i < ob) ? a.actor[l0.off_w + c * l0.ld + i] : 0.f;
        w3s[e] = (i < ob) ? a.actor[l2.off_w + i * l2.ld + c] : 0.f;      // row i = output unit of fc3
    }
    if (threadIdx.x < 64) { b1s[threadIdx.x] = a.actor[l0.off_b + threadIdx.x]; b2s[threadIdx.x] = a.actor[l1.off_b + threadIdx.x]; }
    if (threadIdx.x < kXsLd) {
        const int j = threadIdx.x;
        const float lsj = j < ob ? a.actor[l3.off_w + j] : 0.f;
        const float sd = expf(lsj);      // the per-column terms of the Gaussian log-density, computed once (same operations as per row)
        b3v[j] = j < ob ? a.actor[l2.off_b + j] : 0.f; limv[j] = j < ob ? lim[j] : 0.f;
        varv[j] = __fmul_rn(sd, sd); var2v[j] = __fmul_rn(2.f, __fmul_rn(sd, sd)); logsdv[j] = logf(sd);
    }
    for (int e = threadIdx.x; e < 64 * 16; e += kThreads) {
        const int o = e >> 4, i4 = e & 15;
        const float4 w = *reinterpret_cast<const float4*>(a.actor + l1.off_w + o * l1.ld + 4 * i4);
        float4 h, l;
        split_tf32(w.x, h.x, l.x); split_tf32(w.y, h.y, l.y); split_tf32(w.z, h.z, l.z); split_tf32(w.w, h.w, l.w);
        { const uint32_t base = s0 + kOffW2K + (i4 >> 3) * 16384, off = kmajor_offset(o, i4 & 7); sts128(base + off, h); sts128(base + 8192 + off, l); }
        { const uint32_t base = s0 + kOffW2MN + (o >> 5) * 16384, off = mnmajor_offset(o & 31, i4); sts128(base + off, h); sts128(base + 8192 + off, l); }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int q = warp_id() & 3, hcol = warp_id() >> 2, c0 = 32 * hcol;
    const int r = 32 * q + lane_id();
    const uint32_t my_t = tmem + ((uint32_t)(32 * q) << 16) + c0;
    const uint32_t idesc_fwd = make_idesc_tf32(128, 64, 0, 0), idesc_dx = make_idesc_tf32(128, 64, 0, 1), idesc_dw = make_idesc_tf32(64, 64, 1, 1);
    uint32_t ph_fc2 = 0, ph_dw = 0, ph_dx = 0;
    uint64_t* const m_fc2 = mbar3; uint64_t* const m_dw = mbar3 + 1; uint64_t* const m_dx = mbar3 + 2;
    const float invB = 1.0f / (float)a.b.n_mean;
    const float kLogSqrt2Pi = 0.918938533204672741780329736406f;
    const float clip_lo = 1.f - a.h.epsilon, clip_hi = 1.f + a.h.epsilon;

    float gw2[32], cs_b2[32], gw1[kXsLd], gw3[5];
#pragma unroll
    for (int j = 0; j < 32; ++j) { gw2[j] = 0.f; cs_b2[j] = 0.f; }
#pragma unroll
    for (int m = 0; m < kXsLd; ++m) gw1[m] = 0.f;
#pragma unroll
    for (int m = 0; m < 5; ++m) gw3[m] = 0.f;
    float s_loss = 0.f, s_kl = 0.f, s_dist = 0.f, gb1 = 0.f, colacc = 0.f;      // colacc: threads 0..19 sum d3 (d b3), threads 32..51 the log-scale terms
    const int o1 = threadIdx.x & 63, rq = threadIdx.x >> 6;

    auto load_x_tile = [&](int t0, float* dst) {
        const int trows = min(kTile, nrows - t0);
        const float* X = a.b.x + (r0 + t0) * ldo;
        for (int e = threadIdx.x; e < kTile * (kXsLd / 4); e += kThreads) {
            const int rr = e / (kXsLd / 4), c4 = e % (kXsLd / 4);
            const bool ok = rr < trows && 4 * c4 < ldo;
            cp_async16(dst + rr * kXsLd + 4 * c4, ok ? X + (size_t)rr * ldo + 4 * c4 : a.b.x, ok ? 16 : 0);
        }
        cp_async_commit();
    };
    auto dw1_ffma = [&](const float* xs) {
#pragma unroll 4
        for (int rr = 32 * rq; rr < 32 * rq + 32; ++rr) {
            const float d = dz1s[rr * kDz1Ld + o1];
            gb1 = __fadd_rn(gb1, d);
            const float4* xr = reinterpret_cast<const float4*>(xs + rr * kXsLd);
#pragma unroll
            for (int c = 0; c < kXsLd / 4; ++c) {
                const float4 x4 = xr[c];
                fma2(gw1[4 * c], gw1[4 * c + 1], d, x4.x, x4.y);
                fma2(gw1[4 * c + 2], gw1[4 * c + 3], d, x4.z, x4.w);
            }
        }
    };
    load_x_tile(0, xs_base);
    int tbuf = 0;
    for (int t0 = 0; t0 < nrows; t0 += kTile, tbuf ^= 1) {
        const int trows = min(kTile, nrows - t0);
        float* xs = xs_base + tbuf * (kXsBytes / 4);
        cp_async_wait<0>();
        __syncthreads();
        const bool valid = r < trows;
        const int64_t grow = r0 + t0 + r;
        float old_lp = 0.f, adv = 0.f;
        if (valid && hcol == 0) { old_lp = __ldg(a.b.logp + grow); adv = __ldg(a.b.adv + grow); }
        float h1[32];
        {
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
            float xv[kXsLd];
#pragma unroll
            for (int c = 0; c < kXsLd / 4; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(xs + r * kXsLd + 4 * c);
                xv[4 * c] = v.x; xv[4 * c + 1] = v.y; xv[4 * c + 2] = v.z; xv[4 * c + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < kXsLd; ++i) {
                const float4* wr = reinterpret_cast<const float4*>(w1t + i * 64 + c0);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 w = wr[c];
                    fma2(acc[4 * c], acc[4 * c + 1], xv[i], w.x, w.y);
                    fma2(acc[4 * c + 2], acc[4 * c + 3], xv[i], w.z, w.w);
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) h1[j] = tanhf(__fadd_rn(acc[j], b1s[c0 + j]));
        }
        store_kmajor_row<32>(s0 + kOffBufK, r, c0, h1);
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<false, false>(tmem + 64 * c, s0 + kOffBufK + c * 32768, s0 + kOffBufK + c * 32768 + 16384,
                                        s0 + kOffW2K + c * 16384, s0 + kOffW2K + c * 16384 + 8192, idesc_fwd);
            commit(m_fc2);
        }
        if (t0 > 0) dw1_ffma(xs_base + (tbuf ^ 1) * (kXsBytes / 4));
        mbar_wait(m_fc2, ph_fc2); ph_fc2 ^= 1;
        fence_after_sync();
        __syncthreads();      // every thread is past the previous tile's dW1 loop: bufA | bufB become the head's scratch, its x buffer is free
        if (t0 + kTile < nrows) load_x_tile(t0 + kTile, xs_base + (tbuf ^ 1) * (kXsBytes / 4));
        // ---- h2, the two halves' partial fc3 pre-activations
        float h2[32];
        tmem_row<false, 32>(my_t, h2);
        tmem_row<true, 32>(my_t + 64, h2);
#pragma unroll
        for (int j = 0; j < 32; ++j) h2[j] = tanhf(__fadd_rn(h2[j], b2s[c0 + j]));
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(h2s + r * kDz1Ld + c0 + 4 * c) = make_float4(h2[4 * c], h2[4 * c + 1], h2[4 * c + 2], h2[4 * c + 3]);
        {
            float pre[kXsLd];
#pragma unroll
            for (int j = 0; j < kXsLd; ++j) pre[j] = 0.f;
            for (int j = 0; j < ob; ++j) {
                const float4* wr = reinterpret_cast<const float4*>(w3s + j * 64 + c0);
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 w = wr[c];
                    s = fmaf(h2[4 * c], w.x, s); s = fmaf(h2[4 * c + 1], w.y, s); s = fmaf(h2[4 * c + 2], w.z, s); s = fmaf(h2[4 * c + 3], w.w, s);
                }
                pre[j] = s;
            }
#pragma unroll
            for (int c = 0; c < kXsLd / 4; ++c)
                *reinterpret_cast<float4*>(exch + (hcol * kTile + r) * kXsLd + 4 * c) = make_float4(pre[4 * c], pre[4 * c + 1], pre[4 * c + 2], pre[4 * c + 3]);
        }
        __syncthreads();
        // ---- head, one thread per row: mean, log-prob, ratio / clip (or the A2C form), d3 and the log-scale terms
        if (hcol == 0) {
            float lp = 0.f;
            float tv[kXsLd], dd[kXsLd], av[kXsLd];
#pragma unroll
            for (int c = 0; c < kXsLd / 4; ++c) {      // the row's stored action with 16-byte loads
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && 4 * c < ldo) v = __ldg(reinterpret_cast<const float4*>(a.b.act + grow * ldo + 4 * c));
                av[4 * c] = v.x; av[4 * c + 1] = v.y; av[4 * c + 2] = v.z; av[4 * c + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < kXsLd; ++j) {
                if (j >= ob) { tv[j] = 0.f; dd[j] = 0.f; continue; }
                const float pre = __fadd_rn(__fadd_rn(exch[r * kXsLd + j], exch[(kTile + r) * kXsLd + j]), b3v[j]);
                const float t = tanhf(pre);
                tv[j] = t;
                const float mu = __fmul_rn(t, limv[j]);
                const float d = valid ? __fsub_rn(av[j], mu) : 0.f;
                dd[j] = d;
                lp += __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), var2v[j]), logsdv[j]), kLogSqrt2Pi);
            }
            float dlogp = 0.f;
            if (valid) {
                a.s.newlogp[grow] = lp;
                const float ratio = expf(__fsub_rn(lp, old_lp));
                const float clipped = fminf(fmaxf(ratio, clip_lo), clip_hi);
                const float s1 = __fmul_rn(ratio, adv), s2 = __fmul_rn(clipped, adv);
                s_loss += a.a2c ? -__fmul_rn(old_lp, adv) : -fminf(s1, s2);
                s_kl += __fsub_rn(old_lp, lp);
                const float g = -invB;
                const float tie = (s1 == s2) ? 0.5f * g : 0.f;
                const float g1 = (s1 < s2 ? g : 0.f) + tie, g2 = (s2 < s1 ? g : 0.f) + tie;
                const float in_range = (ratio >= clip_lo && ratio <= clip_hi) ? 1.f : 0.f;
                const float dratio = __fadd_rn(__fmul_rn(g1, adv), __fmul_rn(__fmul_rn(g2, adv), in_range));
                dlogp = __fmul_rn(dratio, ratio);
                if (a.a2c) dlogp = __fmul_rn(g, adv);
            }
            float xnv[kXsLd];
            const bool want_dist = valid && a.h.custom_loss != 0.f;
#pragma unroll
            for (int c = 0; c < kXsLd / 4; ++c) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (want_dist && 4 * c < ldo) v = __ldg(reinterpret_cast<const float4*>(a.b.xn + grow * ldo + 4 * c));
                xnv[4 * c] = v.x; xnv[4 * c + 1] = v.y; xnv[4 * c + 2] = v.z; xnv[4 * c + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < kXsLd; ++j) {
                if (j >= ob) continue;
                const float var = varv[j];
                const float d = dd[j];
                const float dmean = __fmul_rn(dlogp, __fdiv_rn(d, var));
                d3s[r * kXsLd + j] = __fmul_rn(__fmul_rn(dmean, limv[j]), __fsub_rn(1.f, __fmul_rn(tv[j], tv[j])));
                exch[r * kXsLd + j] = __fmul_rn(dlogp, __fsub_rn(__fdiv_rn(__fmul_rn(d, d), var), 1.f));      // d logp / d log_scale term
                if (want_dist) {
                    float la = av[j], ln = xnv[j];
                    if (a.mode == 1) {
                        const float* doff = a.norm + NORM_DOFF * ldo; const float* dsc = a.norm + NORM_DSCALE * ldo;
                        la = __fadd_rn(doff[j], __fmul_rn(la, dsc[j])); ln = __fadd_rn(doff[j], __fmul_rn(ln, dsc[j]));
                    }
                    const float df = __fsub_rn(la, ln);
                    s_dist = fmaf(df, df, s_dist);
                }
            }
            for (int j = ob; j < kXsLd; ++j) { d3s[r * kXsLd + j] = 0.f; exch[r * kXsLd + j] = 0.f; }
        }
        __syncthreads();
        // ---- dz2 = (d3 W3) (1 - h2^2); column sums of d3 / log-scale terms; dW fc3 on FFMA (out unit pairs: column o1, rows j = rq + 4 m)
        float dz2[32];
        {
#pragma unroll
            for (int j = 0; j < 32; ++j) dz2[j] = 0.f;
            for (int j = 0; j < ob; ++j) {
                const float d = d3s[r * kXsLd + j];
                const float4* wr = reinterpret_cast<const float4*>(w3s + j * 64 + c0);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 w = wr[c];
                    fma2(dz2[4 * c], dz2[4 * c + 1], d, w.x, w.y);
                    fma2(dz2[4 * c + 2], dz2[4 * c + 3], d, w.z, w.w);
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) { dz2[j] = __fmul_rn(dz2[j], __fsub_rn(1.f, __fmul_rn(h2[j], h2[j]))); cs_b2[j] += dz2[j]; }
        }
        if (threadIdx.x < kXsLd) { for (int rr = 0; rr < kTile; ++rr) colacc += d3s[rr * kXsLd + threadIdx.x]; }
        else if (threadIdx.x >= 32 && threadIdx.x < 32 + kXsLd) { for (int rr = 0; rr < kTile; ++rr) colacc += exch[rr * kXsLd + threadIdx.x - 32]; }
        for (int rr = 0; rr < kTile; ++rr) {
            const float h = h2s[rr * kDz1Ld + o1];
            const float* dr = d3s + rr * kXsLd + rq;
#pragma unroll
            for (int m = 0; m < 5; ++m) gw3[m] = fmaf(dr[4 * m], h, gw3[m]);
        }
        __syncthreads();      // the head's scratch is read: bufA | bufB take the MN-major images
        if (q < 2) {
            const uint32_t pa = s0 + kOffBufA + (r >> 5) * 16384, pb = s0 + kOffBufB + (r >> 5) * 16384;
            store_both_rows<32>(s0 + kOffBufK, r, pa, pa + 8192, r & 31, c0, dz2);
            store_mnmajor_row<32>(pb, pb + 8192, r & 31, c0, h1);
        } else {
            store_kmajor_row<32>(s0 + kOffBufK, r, c0, dz2);
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<true, true>(tmem + 256 + 64 * c, s0 + kOffBufA + c * 16384, s0 + kOffBufA + c * 16384 + 8192,
                                      s0 + kOffBufB + c * 16384, s0 + kOffBufB + c * 16384 + 8192, idesc_dw);
            commit(m_dw);
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<false, true>(tmem + 128 + 64 * c, s0 + kOffBufK + c * 32768, s0 + kOffBufK + c * 32768 + 16384,
                                       s0 + kOffW2MN + c * 16384, s0 + kOffW2MN + c * 16384 + 8192, idesc_dx);
            commit(m_dx);
        }
        mbar_wait(m_dw, ph_dw); ph_dw ^= 1;
        fence_after_sync();
        if (q >= 2) {
            const int lr = r - 64;
            const uint32_t pa = s0 + kOffBufA + (lr >> 5) * 16384, pb = s0 + kOffBufB + (lr >> 5) * 16384;
            store_mnmajor_row<32>(pa, pa + 8192, lr & 31, c0, dz2);
            store_mnmajor_row<32>(pb, pb + 8192, lr & 31, c0, h1);
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c)
                chunk_mma<true, true>(tmem + 384 + 64 * c, s0 + kOffBufA + c * 16384, s0 + kOffBufA + c * 16384 + 8192,
                                      s0 + kOffBufB + c * 16384, s0 + kOffBufB + c * 16384 + 8192, idesc_dw);
            commit(m_dw);
        }
        mbar_wait(m_dx, ph_dx); ph_dx ^= 1;
        fence_after_sync();
        float dz1[32];
        tmem_row<false, 32>(my_t + 128, dz1);
        tmem_row<true, 32>(my_t + 192, dz1);
#pragma unroll
        for (int j = 0; j < 32; ++j) dz1[j] = __fmul_rn(dz1[j], __fsub_rn(1.f, __fmul_rn(h1[j], h1[j])));
        mbar_wait(m_dw, ph_dw); ph_dw ^= 1;
        fence_after_sync();
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(dz1s + r * kDz1Ld + c0 + 4 * c) = make_float4(dz1[4 * c], dz1[4 * c + 1], dz1[4 * c + 2], dz1[4 * c + 3]);
        {
            float t[32];
            tmem_row<false, 32>(my_t + 256, t);
            tmem_row<true, 32>(my_t + 320, t);
            tmem_row<true, 32>(my_t + 384, t);
            tmem_row<true, 32>(my_t + 448, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) gw2[j] = __fadd_rn(gw2[j], t[j]);
        }
        fence_before_sync();
    }
    __syncthreads();
    dw1_ffma(xs_base + (tbuf ^ 1) * (kXsBytes / 4));

    // ---- the CTA's partial gradient
    if (lane_id() < 16) {
        const int o = 16 * q + lane_id();
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(part + l1.off_w + o * l1.ld + c0 + 4 * c) = make_float4(gw2[4 * c], gw2[4 * c + 1], gw2[4 * c + 2], gw2[4 * c + 3]);
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const int j = rq + 4 * m;
        if (j < ob) part[l2.off_w + j * l2.ld + o1] = gw3[m];
    }
    if (threadIdx.x < ob) part[l2.off_b + threadIdx.x] = colacc;
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + ob) part[l3.off_w + threadIdx.x - 32] = colacc;
    {
        float* w1red = reinterpret_cast<float*>(smem + kOffBufA);
        __syncthreads();
#pragma unroll
        for (int m = 0; m < kXsLd; ++m) w1red[(rq * 64 + o1) * 21 + m] = gw1[m];
        w1red[(rq * 64 + o1) * 21 + kXsLd] = gb1;
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * 21; e += kThreads) {
            const int o = e / 21, i = e % 21;
            const float sum = __fadd_rn(__fadd_rn(__fadd_rn(w1red[(0 * 64 + o) * 21 + i], w1red[(1 * 64 + o) * 21 + i]), w1red[(2 * 64 + o) * 21 + i]),
                                        w1red[(3 * 64 + o) * 21 + i]);
            if (i < ob) part[l0.off_w + o * l0.ld + i] = sum;
            else if (i == kXsLd) part[l0.off_b + o] = sum;
        }
    }
    float* colred = reinterpret_cast<float*>(smem + kOffBufK);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; ++j) colred[(hcol * kTile + r) * 33 + j] = cs_b2[j];
    __syncthreads();
    if (threadIdx.x < 64) {
        const int hh = threadIdx.x >> 5, j = threadIdx.x & 31;
        float s = 0.f;
        for (int rr = 0; rr < kTile; ++rr) s += colred[(hh * kTile + rr) * 33 + j];
        part[l1.off_b + threadIdx.x] = s;
    }
    const float t_loss = block_sum(s_loss, red);
    __syncthreads();
    const float t_kl = block_sum(s_kl, red);
    __syncthreads();
    const float t_dist = block_sum(s_dist, red);
    if (threadIdx.x == 0) { scal[PS_LOSS] = t_loss; scal[PS_KL] = t_kl; scal[PS_DIST] = t_dist; }
    fence_before_sync();
    __syncthreads();
    if (warp_id() == 0) tmem_dealloc<512>(tmem);
}

cudaError_t launch_ppo_actor_grad_tc(const PpoArgs& a, int grid, cudaStream_t s) {
    if (!ppo_critic_tc_supported(a.L.ob, a.L.ldo)) return cudaErrorInvalidValue;
    // no slack for the manual 1024-byte alignment here (230.6 KB + the static 1 KB is all an SM has): the kernel relies on the
    // __align__(1024) of its dynamic shared memory and traps if the base is not aligned
    cudaError_t e = cudaFuncSetAttribute(ppo_actor_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kActSmemBytes);
    if (e != cudaSuccess) return e;
    ppo_actor_grad_tc_kernel<<<grid, kThreads, kActSmemBytes, s>>>(a);
    return cudaGetLastError();
}

}  // namespace spp
