// tcgen05 (5th-generation tensor core) building blocks for sm_100a: TMEM allocation, shared-memory matrix descriptors
// for the SWIZZLE_128B canonical layouts (K-major and MN-major, 32-bit elements), kind::tf32 MMA issue, commit to an
// mbarrier, and TMEM -> register loads.  Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables (same fields as CUTLASS's cute/arch/mma_sm100_desc.hpp, used here as documentation).
//
// fp32 accuracy on tensor cores: every operand x is split as x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi);
// D += hi*hi + hi*lo + lo*hi (three kind::tf32 MMAs into the same TMEM accumulator) leaves a per-product error of
// ~2^-23, the same order as an fp32 FMA chain, which is what the 1e-5 parity bar of this project needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spp {
namespace umma {

constexpr int kAtomBytes = 1024;        // 8 rows x 128 B swizzle atom
constexpr int kChunkK = 32;             // 32 tf32 = 128 B: one swizzle row per tile row and k-chunk
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * 128;   // one operand tile of a k-chunk: 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- shared-memory operand tiles --------------------------------------------------------------------------------
// K-major tile (rows x 32 floats, k contiguous in the source): row r, 16-byte chunk c -> atom r/8, row r%8, chunk c ^ (r%8)
__device__ __forceinline__ uint32_t kmajor_offset(int row, int chunk) {
    return (uint32_t)((row >> 3) * kAtomBytes + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
// MN-major tile (32 k-rows x 128 floats, m/n contiguous in the source).  For 32-bit operands the only MN-major layout the
// tensor core accepts is SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 128 B (32 floats along MN), the 32-byte chunk index of a
// row XORed with the k-row index inside the atom (byte-address bits [5,7) ^= bits [7,9)).  k-row k, 16-byte chunk c of the
// 128-float row -> MN group g = c / 8 (its 32 k-rows are contiguous: 4096 B), atom k / 4 (512 B), row k % 4 (128 B).
__device__ __forceinline__ uint32_t mnmajor_offset(int k, int chunk) {
    const int c8 = chunk & 7, q = c8 >> 1, h = c8 & 1;
    return (uint32_t)((chunk >> 3) * 4096 + (k >> 2) * 512 + (k & 3) * 128 + ((q ^ (k & 3)) << 5) + (h << 4));
}

// 64-bit matrix descriptor: start address [0,14) >>4, leading byte offset [16,30) >>4, stride byte offset [32,46) >>4,
// version (1 on Blackwell) [46,48), layout type [61,64) (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile_addr, int kstep /* 0..3: 8 tf32 = 32 B each */) {
    return make_desc(tile_addr + kstep * 32, 16, kAtomBytes);
}
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile_addr, int kstep /* 0..3: 8 k-rows = two 4-row atoms = 1024 B */) {
    return make_desc(tile_addr + kstep * 1024, 4096, 512, 1 /* SWIZZLE_128B_BASE32B */);
}

// ---- 16-wide k-chunks (64-byte rows): the pipelined GEMM of gemm_umma.cuh stages chunks of 16 tf32 so that FOUR operand slots fit
// the shared memory (copies run three chunks ahead of the tensor core instead of one).
// K-major: SWIZZLE_64B, atoms of 8 rows x 64 B; the 16-byte chunk index (address bits [4,6)) is XORed with address bits [7,9),
// i.e. with (row % 8) / 2.  MN-major: the same SWIZZLE_128B_BASE32B atoms as above, 16 k-rows (4 atoms = 2048 B) per MN group.
__device__ __forceinline__ uint32_t kmajor16_offset(int row, int chunk /* 0..3 */) {
    return (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ uint32_t mnmajor16_offset(int k /* 0..15 */, int chunk) {
    const int c8 = chunk & 7, q = c8 >> 1, h = c8 & 1;
    return (uint32_t)((chunk >> 3) * 2048 + (k >> 2) * 512 + (k & 3) * 128 + ((q ^ (k & 3)) << 5) + (h << 4));
}
__device__ __forceinline__ uint64_t kmajor16_desc(uint32_t tile_addr, int kstep /* 0..1 */) {
    return make_desc(tile_addr + kstep * 32, 16, 512, 4 /* SWIZZLE_64B */);
}
__device__ __forceinline__ uint64_t mnmajor16_desc(uint32_t tile_addr, int kstep /* 0..1 */) {
    return make_desc(tile_addr + kstep * 1024, 2048, 512, 1 /* SWIZZLE_128B_BASE32B */);
}

// 32-bit instruction descriptor for kind::tf32 with fp32 accumulation.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM -------------------------------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {      // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {           // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS));
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes (st.shared / cp.async) -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA / commit / wait ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void commit(uint64_t* mbar) {      // arrives on mbar when all MMAs issued so far have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    const uint32_t addr = smem_u32(mbar);
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive columns (one row per thread)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same without the wait: issue several loads, then tmem_ld_wait() once before the registers are stored
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// split x = hi + lo for the 3-pass scheme: hi = rn_tf32(x), lo = rn_tf32(x - hi), round-to-nearest (ties away) done with
// integer arithmetic on the bit pattern (cvt.rna.tf32.f32 is emulated on sm_100a with an Inf/NaN guard: 4 instructions per
// conversion; operands here are finite, so the guard is dropped: add half an ulp of tf32, clear the 13 low mantissa bits).
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = round_tf32(x);
    lo = round_tf32(__fsub_rn(x, hi));
}

// explicit shared-space 16-byte accesses (generic-pointer float4 accesses compile to LD.E / ST.E when the space is unknown)
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}

}  // namespace umma
}  // namespace spp
