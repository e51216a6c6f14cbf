// Data-parallel SPP-PPO: the per-optimiser-step gradient all-reduce as ONE kernel over NVLink peer memory, fused
// with the reduction of the per-CTA partial gradients that precedes it.  One process per GPU; every rank exports a small exchange
// buffer with cudaIpcGetMemHandle and maps its peers' (spp_ppo_p2p_handle / spp_ppo_p2p_init); NVSwitch gives every GPU a direct
// load / store path to every peer.
//
// Per step (26 KB: the gradient vector and the 8 loss scalars), on every rank:
//   1. thread i sums element i over the CTA slots of the gradient kernel (what ppo_reduce_kernel does) and writes the sum to the
//      rank's OWN exchange slot (epoch parity: two slots);
//   2. the last CTA to finish (atomic ticket) publishes "epoch e is there" by a release store into every PEER's flag word;
//   3. every CTA waits until all peers' flags for epoch e have arrived in its own memory (acquire loads, bounded spin);
//   4. thread i loads element i from every rank's slot through the peer mappings and adds them in RANK ORDER -- the same order on
//      every rank, so all ranks hold bit-identical sums (NCCL's ring / tree orders differ per rank position only in theory, but here
//      it is by construction) -- and writes the result where the Adam kernel reads it.
// A slot written at epoch e was last read by the peers at epoch e - 2; their flag for e - 1 (awaited in step 3 of the previous call)
// was stored after that kernel had finished, so two slots suffice.  No NCCL call, no host synchronisation; ~6 kernels' worth of
// launch latency and one ~30 us collective per step become one ~10 us kernel.
#include "ppo_kernels.cuh"
#include "ppo_p2p.h"

namespace spp {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) ppo_reduce_p2p_kernel(P2pArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = (int)(a.epoch & 1u);
    float* mine = a.peer[a.rank] + (size_t)slot * a.total;
    if (i < a.total) {
        float s;
        if (a.n_part > 0) {      // fused reduction of the gradient kernel's per-CTA slots (fixed order: deterministic)
            s = 0.f;
            if (i < a.n_elems) s = slot_sum(a.part, (size_t)a.part_stride, a.n_part, i);
            else if (i >= a.part_stride) s = slot_sum(a.scal, (size_t)PS_COUNT, a.n_part, i - a.part_stride);
        } else {
            s = a.gbuf[i];       // already reduced (or zeroed: a rank that owns no row of this minibatch)
        }
        mine[i] = s;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(a.ticket, 1u);
        last = (t + 1u == a.ticket_target) ? 1 : 0;
    }
    __syncthreads();
    if (last) {                  // the whole vector of this rank is written: tell every peer (its flag word for this rank and slot)
        __threadfence_system();
        if (threadIdx.x < a.world) st_release_sys(a.peer_flags[threadIdx.x] + slot * kP2pMaxRanks + a.rank, a.epoch);
    }
    // wait for every rank's vector of this epoch (flags arrive in OUR memory); bounded: a lost peer must not hang the GPU
    if (threadIdx.x < a.world) {
        const uint32_t* f = a.peer_flags[a.rank] + slot * kP2pMaxRanks + threadIdx.x;
        long long spins = 0;
        while (ld_acquire_sys(f) != a.epoch) {
            if (++spins > (1ll << 28)) { *a.err = 1; break; }
            __nanosleep(20);
        }
    }
    __syncthreads();
    if (i < a.total) {
        float s = 0.f;
        for (int r = 0; r < a.world; ++r) s += ld_relaxed_sys(a.peer[r] + (size_t)slot * a.total + i);      // rank order: identical on every rank
        a.gbuf[i] = s;
        if (a.step.enabled) {      // the optimiser step of this element, and the step's record slot
            if (i >= a.part_stride) { if (a.step.slog) a.step.slog[i - a.part_stride] = s; }
            else apply_step_element(a.step, i, s);
        }
    }
}

cudaError_t launch_ppo_reduce_p2p(const P2pArgs& a, cudaStream_t s) {
    ppo_reduce_p2p_kernel<<<(a.total + 255) / 256, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace spp
