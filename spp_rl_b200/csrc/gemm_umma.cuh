// Tensor-core GEMM of the fused kernels: C[M x 256] = A . B with M <= 256, on tcgen05 kind::tf32 with the three-pass
// hi/lo split (umma.cuh), accurate to fp32-FMA level.
//
//   A_KM = true : A is [M x K] row-major (contraction contiguous): forward x, dY of a dX product        -> K-major tile
//   A_KM = false: A is [K x M] row-major (contraction = batch): dY of a dW product                      -> MN-major tile
//   B_KM = true : B is [256 x K] row-major (the natural torch weight in a forward)                       -> K-major tile
//   B_KM = false: B is [K x 256] row-major (W^T in a forward, W in dX, the layer input in dW)            -> MN-major tile
//
// Accuracy.  The tensor core adds every MMA's 8-product sum into the fp32 TMEM accumulator with TRUNCATION (measured:
// all-positive data, K = 256, one accumulator: bias -1.6e-6 relative, linear in the number of accumulations; see
// tools/umma_accum_error.py).  A K = 256 product accumulated in place is therefore 6x less accurate than an fp32 FMA chain,
// which the 1e-5 parity bar of this project does not survive through Adam.  Two measures bring it back to fp32 level:
//   (1) the accumulator only ever holds ONE 32-wide k-chunk: chunks alternate between two TMEM accumulators and each
//       finished chunk is drained (tcgen05.ld) and added to a running sum in registers with round-to-nearest fp32 adds
//       while the tensor core works on the next chunk;
//   (2) inside a chunk the small cross terms (lo*hi, hi*lo, 2^-11 of the result) are issued first, while the accumulator
//       is still tiny, so only the 4 hi*hi MMAs of a chunk truncate at full magnitude.
//
// Tiling.  One tile is 128 rows x 256 columns (two tiles for M > 128): TMEM = 2 accumulators x 256 columns.  Shared memory
// = 2 slots x [A_hi 16 KB | A_lo 16 KB | B_hi 32 KB | B_lo 32 KB] in the SWIZZLE_128B layouts of umma.cuh.  Raw fp32 chunks
// travel global -> shared by 16-byte LDGSTS straight into their swizzled position in the hi planes of the free slot (no
// registers in flight: the register file is needed for the 128 running sums); per chunk every thread then splits the 12
// float4 it copied in place (hi rounded in place, lo to the twin plane), one thread issues 12 MMAs of 128 x 256 x 8 and
// commits them to the slot's mbarrier, then everybody drains the previous chunk while the next raw chunk lands.  After the
// last chunk the running sums go through a shared staging buffer to the epilogue functors of gemm_tile.cuh in their
// register mapping (coalesced global traffic).
#pragma once
#include "gemm_tile.cuh"
#include "umma.cuh"

namespace spp {

#ifndef SPP_UMMA_BK
#define SPP_UMMA_BK 32      // k-chunk: 32 tf32 (128-byte rows, two 96 KB slots, copies one chunk ahead) or 16 (64-byte rows, four 48 KB slots, three ahead: measured slower, profiles/r02_update_bk16_experiment.md)
#endif
#ifndef SPP_UMMA_DRAIN_EVERY
#define SPP_UMMA_DRAIN_EVERY 1      // k-chunks one TMEM accumulator collects before it is drained (1: the accurate default; 2: measured, see profiles/)
#endif
constexpr int kDrainEvery = SPP_UMMA_DRAIN_EVERY;
constexpr int kBK = SPP_UMMA_BK;
static_assert(kBK == 16 || kBK == 32, "k-chunk of the tcgen05 pipeline");
constexpr int kKSteps = kBK / 8;                           // MMAs of k = 8 per pass and chunk
constexpr int kUmmaAPlane = 128 * kBK * 4;                 // A_hi / A_lo: 128 rows x kBK tf32
constexpr uint32_t kUmmaSinglePass = 8;   // UmmaCtx::dbg bit: reduced-precision variant (one tf32 pass, whole K in one TMEM accumulator)
constexpr int kUmmaBPlane = 256 * kBK * 4;                 // B_hi / B_lo: 256 columns x kBK tf32
constexpr int kUmmaSlotBytes = 2 * kUmmaAPlane + 2 * kUmmaBPlane;      // 48 KB (96 KB)
constexpr int kUmmaSlots = 64 / kBK;                                   // 4 (2)
constexpr int kUmmaAhead = kUmmaSlots - 1;                             // chunks the copies run ahead of the MMAs
constexpr int kUmmaSmemBytes = kUmmaSlots * kUmmaSlotBytes;            // 192 KB

namespace umma {   // layouts of one operand chunk for the configured kBK
__device__ __forceinline__ uint32_t km_offset(int row, int chunk) { return kBK == 32 ? kmajor_offset(row, chunk) : kmajor16_offset(row, chunk); }
__device__ __forceinline__ uint32_t mn_offset(int k, int chunk) { return kBK == 32 ? mnmajor_offset(k, chunk) : mnmajor16_offset(k, chunk); }
__device__ __forceinline__ uint64_t km_desc(uint32_t a, int ks) { return kBK == 32 ? kmajor_desc(a, ks) : kmajor16_desc(a, ks); }
__device__ __forceinline__ uint64_t mn_desc(uint32_t a, int ks) { return kBK == 32 ? mnmajor_desc(a, ks) : mnmajor16_desc(a, ks); }
}  // namespace umma
constexpr int kUmmaTmemCols = 512;
constexpr int kStagePitch = 132;                           // floats; conflict-free accumulator staging
constexpr int kStageBlockFloats = 128 * kStagePitch;       // one 128 x 128 block
static_assert(2 * kStageBlockFloats * 4 <= kUmmaSmemBytes, "accumulator staging aliases the operand slots");

// Pipeline state of a CTA, all of it in SHARED memory (a struct in the kernel's frame would sit in local memory, which the 213 KB
// shared-memory carve-out leaves almost no L1 for): the GEMM entry reads it once, thread 0 writes the parities back.
struct UmmaCtx {
    unsigned char* smem;        // 1024-byte aligned, kUmmaSmemBytes
    uint64_t* mbar;             // [kUmmaSlots]: "the MMAs of the chunk in this slot have retired"
    uint32_t tmem;              // TMEM base (512 columns)
    uint32_t phase_bits;        // per-slot mbarrier parity the threads wait for next (identical in every thread)
    uint32_t dbg;               // timing experiments only (self-test): 1 = no MMAs, 2 = no operand staging, 4 = no epilogue
    uint32_t zraw;              // shared-space address of a 32 KB landing zone for the NEXT raw B chunk (umma_mainloop_z), or 0: none
};

// Per-thread state of one operand's global -> register -> shared path.  ROWS = 128 (A) or 256 (B); NP float4 per thread
// and 32-wide k-chunk.  Source pointer and swizzled destination are affine in the chunk index p.
template <bool KM, int ROWS>
struct UmmaOperand {
    static constexpr int NP = ROWS * (kBK / 4) / kThreads;      // float4 per thread and chunk: 4 (8 with 32-wide chunks)
    static constexpr int QPR = kBK / 4;                         // K-major: float4 per row and chunk
    static constexpr int RPP = kThreads / QPR;                  // K-major: rows covered by the CTA per p (64 or 32)
    static constexpr int CPR = ROWS / 4;                        // MN-major: float4 per k-row (32 or 64)
    static constexpr int KSTEP = KM ? 0 : kThreads / CPR;       // MN-major: k-rows between consecutive p (8 or 4)
    static constexpr uint32_t DSTEP = KM ? 4096u : (uint32_t)(KSTEP / 4) * 512u;      // K-major: RPP rows = 4096 B in both layouts
    const float* src0;
    uint32_t sstep;             // floats between consecutive p   (32-bit: an operand is far smaller than 2^32 floats, and every
    uint32_t kstride;           // floats per unit of k            register counts next to the 128 running sums of the main loop)
    uint32_t dst0;
    int kofs0;                  // k offset of chunk p inside the k-chunk: kofs0 + p * KSTEP
    int np_ok;                  // chunks p < np_ok are inside the operand (rows / columns tail)
    // r0: first row (column) of the tile inside the operand, R: rows (columns) of the whole operand
    __device__ __forceinline__ void init(const float* __restrict__ G, int ld, int R, int r0) {
        if constexpr (KM) {      // quarter-warps write 128 contiguous (swizzled) bytes: 8 distinct 16-byte positions
            int q, row;
            if constexpr (kBK == 32) { q = (threadIdx.x >> 3) & 7; row = (threadIdx.x >> 6) * 8 + (threadIdx.x & 7); }
            else { q = threadIdx.x & 3; row = threadIdx.x >> 2; }           // + RPP p
            kstride = 1; kofs0 = 4 * q;
            src0 = G + (size_t)(r0 + row) * ld + 4 * q;
            sstep = (uint32_t)RPP * (uint32_t)ld;
            dst0 = umma::km_offset(row, q);
            const int left = R - r0 - row;                                   // rows row + RPP p < R - r0
            np_ok = left <= 0 ? 0 : min(NP, (left + RPP - 1) / RPP);
        } else {                 // a warp covers 512 contiguous bytes of one k-row
            const int chunk = threadIdx.x % CPR, kb = threadIdx.x / CPR;     // k = kb + KSTEP p
            kstride = (uint32_t)ld; kofs0 = kb;
            src0 = G + (size_t)kb * ld + r0 + 4 * chunk;
            sstep = (uint32_t)KSTEP * (uint32_t)ld;
            dst0 = umma::mn_offset(kb, chunk);
            np_ok = (r0 + 4 * chunk < R) ? NP : 0;
        }
    }
    // raw fp32 chunk, global -> its swizzled position in the hi plane (16-byte LDGSTS, zero-filled outside the operand)
    __device__ __forceinline__ void issue(int k0, int K, uint32_t hi) const {
        const float* s = src0 + (uint32_t)k0 * kstride;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const bool ok = (p < np_ok) && (k0 + kofs0 + p * KSTEP < K);
            const float* g = ok ? s + (uint32_t)p * sstep : src0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(hi + dst0 + p * DSTEP), "l"(g), "r"(ok ? 16 : 0));
        }
    }
    // in-place split of the chunks this thread copied: hi plane <- rn_tf32(x), lo plane <- rn_tf32(x - hi)
    __device__ __forceinline__ void split(uint32_t hi, uint32_t lo) const { split_from(hi, hi, lo); }
    // raw words landing zone -> hi plane, unsplit (the single-pass variant's operands are the raw fp32 words)
    __device__ __forceinline__ void move_from(uint32_t raw, uint32_t hi) const {
#pragma unroll
        for (int p0 = 0; p0 < NP; p0 += 4) {
            float4 r[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) r[p] = umma::lds128(raw + dst0 + (p0 + p) * DSTEP);
#pragma unroll
            for (int p = 0; p < 4; ++p) umma::sts128(hi + dst0 + (p0 + p) * DSTEP, r[p]);
        }
    }
    // the same with the raw words at `raw` (same offsets: the landing zone of umma_mainloop_z, or the hi plane itself)
    __device__ __forceinline__ void split_from(uint32_t raw, uint32_t hi, uint32_t lo) const {
        // four chunks at a time: the 128 running sums of the main loop leave ~100 registers for everything else
#pragma unroll
        for (int p0 = 0; p0 < NP; p0 += 4) {
            float4 r[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) r[p] = umma::lds128(raw + dst0 + (p0 + p) * DSTEP);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float4 h, l;
                umma::split_tf32(r[p].x, h.x, l.x); umma::split_tf32(r[p].y, h.y, l.y);
                umma::split_tf32(r[p].z, h.z, l.z); umma::split_tf32(r[p].w, h.w, l.w);
                umma::sts128(hi + dst0 + (p0 + p) * DSTEP, h);
                umma::sts128(lo + dst0 + (p0 + p) * DSTEP, l);
            }
        }
    }
};

// The A operand with its storage order as a RUNTIME flag (same affine scheme): the main loop is then ONE function for forward / dX
// (K-major A) and dW (MN-major A) products.  With two template instances ptxas gave only one of them a full register budget and
// kept the other's 128 running sums in local memory (~1.1 KB of spills, LDL / STL in every drain).
struct UmmaOperandA {
    static constexpr int NP = 128 * (kBK / 4) / kThreads;       // 2 (4)
    static constexpr int RPP = kThreads / (kBK / 4);            // K-major: rows per p
    const float* src0;
    uint32_t sstep, kstride, dst0, dstep;
    int kofs0, kstep, np_ok;
    __device__ __forceinline__ void init(bool km, const float* __restrict__ G, int ld, int R, int r0) {
        if (km) {
            int q, row;
            if constexpr (kBK == 32) { q = (threadIdx.x >> 3) & 7; row = (threadIdx.x >> 6) * 8 + (threadIdx.x & 7); }
            else { q = threadIdx.x & 3; row = threadIdx.x >> 2; }           // + RPP p
            kstride = 1; kofs0 = 4 * q; kstep = 0; dstep = 4096u;
            src0 = G + (size_t)(r0 + row) * ld + 4 * q;
            sstep = (uint32_t)RPP * (uint32_t)ld;
            dst0 = umma::km_offset(row, q);
            const int left = R - r0 - row;
            np_ok = left <= 0 ? 0 : min(NP, (left + RPP - 1) / RPP);
        } else {
            const int chunk = threadIdx.x % 32, kb = threadIdx.x / 32;       // k = kb + 8 p
            kstride = (uint32_t)ld; kofs0 = kb; kstep = 8; dstep = 1024u;
            src0 = G + (size_t)kb * ld + r0 + 4 * chunk;
            sstep = 8u * (uint32_t)ld;
            dst0 = umma::mn_offset(kb, chunk);
            np_ok = (r0 + 4 * chunk < R) ? NP : 0;
        }
    }
    __device__ __forceinline__ void issue(int k0, int K, uint32_t hi) const {
        const float* s = src0 + (uint32_t)k0 * kstride;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const bool ok = (p < np_ok) && (k0 + kofs0 + p * kstep < K);
            const float* g = ok ? s + (uint32_t)p * sstep : src0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(hi + dst0 + p * dstep), "l"(g), "r"(ok ? 16 : 0));
        }
    }
    __device__ __forceinline__ void split(uint32_t hi, uint32_t lo) const {
        float4 r[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) r[p] = umma::lds128(hi + dst0 + p * dstep);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float4 h, l;
            umma::split_tf32(r[p].x, h.x, l.x); umma::split_tf32(r[p].y, h.y, l.y);
            umma::split_tf32(r[p].z, h.z, l.z); umma::split_tf32(r[p].w, h.w, l.w);
            umma::sts128(hi + dst0 + p * dstep, h);
            umma::sts128(lo + dst0 + p * dstep, l);
        }
    }
};

// drain one finished chunk: this thread's row (TMEM lane) x 128 columns, added to the running sums
template <bool FIRST>
__device__ __forceinline__ void umma_drain(uint32_t taddr, float (&sum)[128]) {
#pragma unroll
#ifndef SPP_DRAIN_STEP
#define SPP_DRAIN_STEP 2      // TMEM loads in flight per wait: 4 costs 16 of the 128 running sums their registers (local-memory spills)
#endif
    for (int cb = 0; cb < 8; cb += SPP_DRAIN_STEP) {
        float v[SPP_DRAIN_STEP][16];
#pragma unroll
        for (int i = 0; i < SPP_DRAIN_STEP; ++i) umma::tmem_ld16_nowait(taddr + (cb + i) * 16, v[i]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < SPP_DRAIN_STEP; ++i)
#pragma unroll
            for (int e = 0; e < 16; ++e) sum[(cb + i) * 16 + e] = FIRST ? v[i][e] : __fadd_rn(sum[(cb + i) * 16 + e], v[i][e]);
    }
}

// staged accumulator (two 128 x 128 blocks) -> registers in the FFMA tiles' mapping -> epilogue functor.  Its own function so
// that the epilogue's registers (Adam: 4 x 4 float4 of loads in flight) are allocated independently of the main loop's.
template <bool A_KM, class Epi>
__device__ __noinline__ void umma_epilogue(uint32_t stage_s, float* stage, int m0, int M, const Epi& epi_ref) {
    Epi epi = epi_ref;      // functor fields in registers: through the reference every global store forces their reload from local memory
    const int tx = threadIdx.x % BigTile::TX, ty = threadIdx.x / BigTile::TX;
#pragma unroll 1
    for (int nh = 0; nh < 2; ++nh) {
        float acc[BigTile::MI][BigTile::NJ];
#pragma unroll
        for (int i = 0; i < BigTile::MI; ++i) {
            const int r = row_of<BigTile, A_KM>(i, ty);
#pragma unroll
            for (int g = 0; g < BigTile::NJ / 4; ++g) {
                const float4 v = umma::lds128(stage_s + 4 * (nh * kStageBlockFloats + r * kStagePitch + col_of<BigTile>(4 * g, tx)));
                acc[i][4 * g] = v.x; acc[i][4 * g + 1] = v.y; acc[i][4 * g + 2] = v.z; acc[i][4 * g + 3] = v.w;
            }
        }
        __syncthreads();      // block nh of the staging is free (the epilogue may use it for column sums)
        epi.template apply<BigTile, A_KM>(acc, m0, nh * 128, M, 256, stage + nh * kStageBlockFloats);
    }
}

// Main loop of one 128 x 256 tile: leaves the fp32 result in the shared staging buffer (two 128 x 128 blocks of pitch
// kStagePitch at the start of the operand slots).  A separate function (independent of the epilogue type) so that its
// register allocation -- 128 running sums + 48 registers of operands in flight -- is not disturbed by the epilogue's.
template <bool B_KM>
__device__ __noinline__ uint32_t umma_mainloop(bool A_KM, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int M,
                                               int K, int m0, unsigned char* smem, uint64_t* mbar, uint32_t tmem, uint32_t phase_in, uint32_t dbg) {
    using namespace umma;
    constexpr int N = 256;
    const int nchunks = (K + kBK - 1) / kBK;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t idesc = make_idesc_tf32(128, N, A_KM ? 0 : 1, B_KM ? 0 : 1);
    const int wq = warp_id() & 3, chalf = warp_id() >> 2;
    const uint32_t my_tmem = tmem + ((uint32_t)(32 * wq) << 16) + chalf * 128;      // this thread's lane / column half
    const bool fast = (dbg & kUmmaSinglePass) != 0;
    uint32_t phase_bits = phase_in;      // everything by value: a by-reference argument would pin the caller's state in local memory
    UmmaOperand<B_KM, 256> lb;
    lb.init(B, ldb, N, 0);
    UmmaOperandA la;
    la.init(A_KM, A, lda, M, m0);
    float sum[128];
    // raw chunks 0 .. kUmmaAhead-1 -> their slots; one cp.async group per chunk (empty past the end), so that "all but the newest
    // kUmmaAhead-1 groups have landed" means "chunk c is in shared memory" at the top of iteration c
#pragma unroll
    for (int c = 0; c < kUmmaAhead; ++c) {
        if (c < nchunks && !(dbg & 2)) {
            const uint32_t sl = smem0 + c * kUmmaSlotBytes;
            la.issue(kBK * c, K, sl); lb.issue(kBK * c, K, sl + 2 * kUmmaAPlane);
        }
        cp_async_commit();
    }
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const int slot = c % kUmmaSlots;
        const uint32_t ah = smem0 + slot * kUmmaSlotBytes, al = ah + kUmmaAPlane, bh = al + kUmmaAPlane, bl = bh + kUmmaBPlane;
        cp_async_wait<kUmmaAhead - 1>();       // this thread's copies of chunk c have landed
        if (!(dbg & (2 | kUmmaSinglePass))) {
            la.split(ah, al);
            lb.split(bh, bl);
        }
        fence_proxy_async();
        fence_before_sync();      // orders this thread's drain of chunk c - 2 (same accumulator) before the MMAs below
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
            const uint32_t d = tmem + (fast ? 0 : ((c / kDrainEvery) & 1) * N);
            const bool fresh = (c % kDrainEvery) == 0;      // first chunk of its accumulator group
            const int nks = (dbg & 1) ? 0 : min(kKSteps, (K - kBK * c + 7) / 8);
            if (fast) {     // reduced-precision variant: the raw fp32 words are the tf32 operands, the whole K accumulates in TMEM
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t dah = A_KM ? km_desc(ah, ks) : mn_desc(ah, ks);
                    const uint64_t dbh = B_KM ? km_desc(bh, ks) : mn_desc(bh, ks);
                    mma_tf32(d, dah, dbh, idesc, (c | ks) ? 1u : 0u);
                }
            } else {
            for (int ks = 0; ks < nks; ++ks) {      // cross terms first: the accumulator is still tiny
                const uint64_t dah = A_KM ? km_desc(ah, ks) : mn_desc(ah, ks);
                const uint64_t dal = A_KM ? km_desc(al, ks) : mn_desc(al, ks);
                const uint64_t dbh = B_KM ? km_desc(bh, ks) : mn_desc(bh, ks);
                const uint64_t dbl = B_KM ? km_desc(bl, ks) : mn_desc(bl, ks);
                mma_tf32(d, dal, dbh, idesc, (ks || !fresh) ? 1u : 0u);
                mma_tf32(d, dah, dbl, idesc, 1u);
            }
            for (int ks = 0; ks < nks; ++ks) {
                const uint64_t dah = A_KM ? km_desc(ah, ks) : mn_desc(ah, ks);
                const uint64_t dbh = B_KM ? km_desc(bh, ks) : mn_desc(bh, ks);
                mma_tf32(d, dah, dbh, idesc, 1u);
            }
            }
            commit(mbar + slot);
        }
        const int ps = (c + kUmmaSlots - 1) % kUmmaSlots;      // slot of chunk c - 1 == slot of chunk c + kUmmaAhead
        if (c >= 1) {       // chunk c - 1 has retired: its slot is free, its accumulator complete
            mbar_wait(mbar + ps, (phase_bits >> ps) & 1u);
            phase_bits ^= (1u << ps);
            fence_after_sync();
        }
        if (c + kUmmaAhead < nchunks && !(dbg & 2)) {      // raw chunk c + kUmmaAhead -> the free slot (never used before when c == 0)
            const uint32_t nh = smem0 + ps * kUmmaSlotBytes;
            la.issue(kBK * (c + kUmmaAhead), K, nh);
            lb.issue(kBK * (c + kUmmaAhead), K, nh + 2 * kUmmaAPlane);
        }
        cp_async_commit();
        if (c >= 1 && !fast && (c % kDrainEvery) == 0) {       // the group that ended with chunk c - 1 is complete: drain it while the tensor core works on chunk c
            if (c == kDrainEvery) umma_drain<true>(my_tmem + (((c - 1) / kDrainEvery) & 1) * N, sum);
            else umma_drain<false>(my_tmem + (((c - 1) / kDrainEvery) & 1) * N, sum);
        }
    }
    {   // last chunk
        const int ps = (nchunks - 1) % kUmmaSlots, acc = ((nchunks - 1) / kDrainEvery) & 1;
        mbar_wait(mbar + ps, (phase_bits >> ps) & 1u);
        phase_bits ^= (1u << ps);
        fence_after_sync();
        if (nchunks <= kDrainEvery || fast) umma_drain<true>(my_tmem + (fast ? 0 : acc * N), sum);
        else umma_drain<false>(my_tmem + acc * N, sum);
        fence_before_sync();
    }
    cp_async_wait<0>();      // (only empty groups are left)
    {   // running sums -> shared staging; every MMA that read the slots has retired
        const int srow = 32 * wq + lane_id();
        const uint32_t base = smem0 + 4 * (chalf * kStageBlockFloats + srow * kStagePitch);
#pragma unroll
        for (int q = 0; q < 32; ++q) sts128(base + 16 * q, make_float4(sum[4 * q], sum[4 * q + 1], sum[4 * q + 2], sum[4 * q + 3]));
    }
    __syncthreads();
    return phase_bits;
}


// The same main loop with a LANDING ZONE for the raw B operand (UmmaCtx::zraw: 32 KB outside the operand slots).  With two slots the
// loop above can only request the raw chunk c + 1 once the MMAs of chunk c - 1 have retired, i.e. at the start of iteration c, so its
// split always runs after the copy latency, while the tensor core is idle (staging, MMAs and the drain add up without overlap: 37.6 us
// per 256-cubed product = 15.4 + 12.5 + 9.3, profiles/r02_update_bk16_experiment.md).  Here raw B (two thirds of the staging work)
// is requested a whole iteration earlier into the landing zone and split from there into the free slot WHILE the MMAs of chunk c
// run; raw A (16 KB) still lands in place and is split after the drain, by which time it has arrived.  Every thread re-reads only the
// 16-byte words it copied itself, so the landing zone needs no barrier of its own.  Same MMAs, same drains, same order of additions
// as umma_mainloop: the results are bitwise the same.
// The reduced-precision variant (dbg & kUmmaSinglePass: raw words as operands, one pass, whole K in one accumulator) runs through
// the same loop: raw B is moved from the landing zone to the slot unsplit, nothing is drained before the end.
#ifdef SPP_UMMA_LANDING_ZONE
template <bool B_KM>
__device__ __noinline__ uint32_t umma_mainloop_z(bool A_KM, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int M,
                                                 int K, int m0, unsigned char* smem, uint64_t* mbar, uint32_t tmem, uint32_t phase_in, uint32_t dbg,
                                                 uint32_t zraw) {
    using namespace umma;
    static_assert(kBK == 32 && kDrainEvery == 1 && kUmmaSlots == 2, "the landing-zone loop is written for two 32-wide slots");
    constexpr int N = 256;
    const int nchunks = (K + kBK - 1) / kBK;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t idesc = make_idesc_tf32(128, N, A_KM ? 0 : 1, B_KM ? 0 : 1);
    const int wq = warp_id() & 3, chalf = warp_id() >> 2;
    const uint32_t my_tmem = tmem + ((uint32_t)(32 * wq) << 16) + chalf * 128;      // this thread's lane / column half
    const bool fast = (dbg & kUmmaSinglePass) != 0;
    uint32_t phase_bits = phase_in;
    UmmaOperand<B_KM, 256> lb;
    lb.init(B, ldb, N, 0);
    UmmaOperandA la;
    la.init(A_KM, A, lda, M, m0);
    float sum[128];
    // chunk 0: raw A -> hi plane of slot 0, raw B -> landing zone; both split into slot 0; then raw B of chunk 1 -> landing zone
    la.issue(0, K, smem0);
    lb.issue(0, K, zraw);
    cp_async_commit();
    cp_async_wait<0>();
    if (fast) lb.move_from(zraw, smem0 + 2 * kUmmaAPlane);
    else {
        la.split(smem0, smem0 + kUmmaAPlane);
        lb.split_from(zraw, smem0 + 2 * kUmmaAPlane, smem0 + 2 * kUmmaAPlane + kUmmaBPlane);
    }
    if (1 < nchunks) lb.issue(kBK, K, zraw);
    cp_async_commit();                      // group [B(1)] (empty when there is no chunk 1)
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const int slot = c & 1, ps = slot ^ 1;
        const uint32_t ah = smem0 + slot * kUmmaSlotBytes, al = ah + kUmmaAPlane, bh = al + kUmmaAPlane, bl = bh + kUmmaBPlane;
        fence_proxy_async();      // the split planes of chunk c (st.shared) -> visible to the tensor core
        fence_before_sync();      // orders this thread's drain of chunk c - 2 (same accumulator) before the MMAs below
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
            const uint32_t d = tmem + (fast ? 0 : (c & 1) * N);
            const int nks = min(kKSteps, (K - kBK * c + 7) / 8);
            if (fast) {     // reduced-precision variant: the raw fp32 words are the tf32 operands, the whole K accumulates in TMEM
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t dah = A_KM ? km_desc(ah, ks) : mn_desc(ah, ks);
                    const uint64_t dbh = B_KM ? km_desc(bh, ks) : mn_desc(bh, ks);
                    mma_tf32(d, dah, dbh, idesc, (c | ks) ? 1u : 0u);
                }
            } else {
                for (int ks = 0; ks < nks; ++ks) {      // cross terms first: the accumulator is still tiny
                    const uint64_t dah = A_KM ? km_desc(ah, ks) : mn_desc(ah, ks);
                    const uint64_t dal = A_KM ? km_desc(al, ks) : mn_desc(al, ks);
                    const uint64_t dbh = B_KM ? km_desc(bh, ks) : mn_desc(bh, ks);
                    const uint64_t dbl = B_KM ? km_desc(bl, ks) : mn_desc(bl, ks);
                    mma_tf32(d, dal, dbh, idesc, ks ? 1u : 0u);
                    mma_tf32(d, dah, dbl, idesc, 1u);
                }
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t dah = A_KM ? km_desc(ah, ks) : mn_desc(ah, ks);
                    const uint64_t dbh = B_KM ? km_desc(bh, ks) : mn_desc(bh, ks);
                    mma_tf32(d, dah, dbh, idesc, 1u);
                }
            }
            commit(mbar + slot);
        }
        if (c >= 1) {       // chunk c - 1 has retired: its slot is free, its accumulator complete
            mbar_wait(mbar + ps, (phase_bits >> ps) & 1u);
            phase_bits ^= (1u << ps);
            fence_after_sync();
        }
        const bool more = c + 1 < nchunks;
        const uint32_t nah = smem0 + ps * kUmmaSlotBytes, nal = nah + kUmmaAPlane, nbh = nal + kUmmaAPlane, nbl = nbh + kUmmaBPlane;
        if (more) {
            la.issue(kBK * (c + 1), K, nah);      // raw A of chunk c + 1 -> its place in the free slot
            cp_async_commit();                    // group [A(c + 1)]
            cp_async_wait<1>();                   // the older group [B(c + 1)]: this thread's raw B words are in the landing zone
            if (fast) lb.move_from(zraw, nbh);
            else lb.split_from(zraw, nbh, nbl);   // ... and go to the free slot while the tensor core works on chunk c
            if (c + 2 < nchunks) lb.issue(kBK * (c + 2), K, zraw);
            cp_async_commit();                    // group [B(c + 2)] (possibly empty)
        }
        if (c >= 1 && !fast) {       // drain chunk c - 1
            if (c == 1) umma_drain<true>(my_tmem + ps * N, sum);
            else umma_drain<false>(my_tmem + ps * N, sum);
        }
        if (more) {
            cp_async_wait<1>();                   // all but the newest group: raw A of chunk c + 1 has landed
            if (!fast) la.split(nah, nal);
        }
    }
    {   // last chunk
        const int ps = (nchunks - 1) & 1;
        mbar_wait(mbar + ps, (phase_bits >> ps) & 1u);
        phase_bits ^= (1u << ps);
        fence_after_sync();
        if (nchunks <= 1 || fast) umma_drain<true>(my_tmem + (fast ? 0 : ps * N), sum);
        else umma_drain<false>(my_tmem + ps * N, sum);
        fence_before_sync();
    }
    cp_async_wait<0>();      // (only empty groups are left)
    {   // running sums -> shared staging; every MMA that read the slots has retired
        const int srow = 32 * wq + lane_id();
        const uint32_t base = smem0 + 4 * (chalf * kStageBlockFloats + srow * kStagePitch);
#pragma unroll
        for (int q = 0; q < 32; ++q) sts128(base + 16 * q, make_float4(sum[4 * q], sum[4 * q + 1], sum[4 * q + 2], sum[4 * q + 3]));
    }
    __syncthreads();
    return phase_bits;
}
#endif      // SPP_UMMA_LANDING_ZONE


// The wrapper is a real call in the update kernel: the main loop and the epilogues get their register budget from the live set
// of their caller (ptxas allocates across the call graph), and inlined into the 35 k-instruction kernel body that live set cost the
// epilogues ~0.5 KB and the body 2.5 KB of spills through a ~15 KB L1 (+7.7 % updates/s as a call).  The rollout kernel has
// two wide products per step and a small body: there the call itself costs more (-5 %), so it keeps the wrapper inline.
#ifdef SPP_UMMA_WRAPPER_INLINE
#define SPP_UMMA_WRAPPER_ATTR __forceinline__
#else
#define SPP_UMMA_WRAPPER_ATTR __noinline__
#endif
template <bool A_KM, bool B_KM, class Epi>
__device__ SPP_UMMA_WRAPPER_ATTR void gemm256_umma(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int M, int K,
                                             UmmaCtx& u, Epi& epi) {
    const int mtiles = (M + 127) / 128;
    unsigned char* smem = u.smem; uint64_t* mbar = u.mbar;
    const uint32_t tmem = u.tmem, dbg = u.dbg, zraw = u.zraw;
    uint32_t phase = u.phase_bits;
    for (int mt = 0; mt < mtiles; ++mt) {
#ifdef SPP_UMMA_LANDING_ZONE      // (one main loop per kernel: a second one in the call graph costs the kernel body its registers, 1.4 -> 9.3 KB of spills)
        phase = umma_mainloop_z<B_KM>(A_KM, A, lda, B, ldb, M, K, mt * 128, smem, mbar, tmem, phase, dbg, zraw);
#else
        phase = umma_mainloop<B_KM>(A_KM, A, lda, B, ldb, M, K, mt * 128, smem, mbar, tmem, phase, dbg);
        (void)zraw;
#endif
        if (!(dbg & 4)) umma_epilogue<A_KM, Epi>(umma::smem_u32(smem), reinterpret_cast<float*>(smem), mt * 128, M, epi);
        if (threadIdx.x == 0) u.phase_bits = phase;
        __syncthreads();          // the staging aliases the operand slots of the next tile / GEMM; publishes the parities
    }
}

}  // namespace spp
