// p2p.cuh — the all-gather of per-shard top-K lists as the library's own kernels over NVLink peer
// memory (node-axis sharding, world > 1; DESIGN.md §7, SURVEY.md §8e "Collective" row).
//
// Every rank owns one exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by
// its peers; inside one process the raw pointer is shared):
//     data  [2 parities][world sources][rows_cap][KS] u64
//     flags [2 parities][world sources] u64 (16 words apart)
// Phase s of the SPMD call sequence (a wave's lists, or the restricted reselect of its exclusive
// steps) uses parity s & 1:
//   push  fused into the producer (k_shard_select, select_fast.cuh): every CTA stores its rows into slot
//         [parity][g] of EVERY rank's buffer (peer stores over NVLink), fences system-wide, and the last
//         CTA of the grid publishes flags[parity][g] = seq on every rank (st.release.sys);
//   wait  fused into the consumer (k_merge / k_greedy, select.cuh): lane r of the CTA's first warp spins
//         (ld.acquire.sys) until flags[parity][r] >= seq, bounded by a clock64 timeout that raises *err
//         instead of hanging the GPU (p2p_wait_cta below).
// Two parities suffice: a rank pushes phase s + 2 only after its own consumer of phase s + 1 ran, which
// needed every peer's push of s + 1, which those peers issued after their consumers of phase s.
// No NCCL call, no host round trip on the step path.
#pragma once
#include "kernels.cuh"

namespace rbgtopo {

constexpr int P2P_MAX_WORLD = 16;
constexpr int P2P_FLAG_STRIDE = 16;  // u64 words between flags (separate 128-byte lines)

struct P2PDev {
  int world, rank;
  unsigned long long* peer[P2P_MAX_WORLD];  // every rank's exchange buffer, as mapped here (peer[rank] = own)
  long long slot_stride;                    // u64 elements per [parity][source] slot = rows_cap * KS
  long long flags_off;                      // u64 offset of the flag block inside a buffer
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Prologue of the consuming kernel: lane r < world of warp 0 spins on source r, then the CTA proceeds.
// world <= 1: nothing to wait for.
struct P2PWait {
  const unsigned long long* flags;  // own flag block + parity * world * P2P_FLAG_STRIDE
  unsigned long long seq;
  long long timeout_cycles;
  int* err;
  int world;
};
__device__ __forceinline__ void p2p_wait_cta(const P2PWait& w) {
  if (w.world <= 1) return;
  if ((int)threadIdx.x < w.world) {
    const unsigned long long* f = w.flags + (long long)threadIdx.x * P2P_FLAG_STRIDE;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(f) < w.seq) {
      if (clock64() - t0 > w.timeout_cycles) {
        *w.err = 1;
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
}

}  // namespace rbgtopo
