// p2p.cuh — the all-gather of per-shard top-K lists as the library's own kernels over NVLink peer
// memory (node-axis sharding, world > 1; DESIGN.md §7, SURVEY.md §8e "Collective" row).
//
// Every rank owns one exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by
// its peers; inside one process the raw pointer is shared):
//     data  [2 parities][world sources][rows_cap][KS] u64
//     flags [2 parities][world sources] u64 (16 words apart)
// Phase s of the SPMD call sequence (a wave's lists, or the restricted reselect of its exclusive
// steps) uses parity s & 1:
//   k_p2p_push  rank g copies its rows into slot [parity][g] of EVERY rank's buffer (peer stores over
//               NVLink, 16-byte vectors), fences system-wide, and the last CTA publishes
//               flags[parity][g] = seq on every rank (st.release.sys);
//   k_p2p_wait  one warp: lane r spins (ld.acquire.sys) until flags[parity][r] >= seq, bounded by a
//               clock64 timeout that raises *err instead of hanging the GPU; the consumer (k_merge /
//               k_greedy) is simply the next kernel of the stream.
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

// src: this rank's rows of the phase (n_u64 = rows * KS elements, 16-byte aligned); done: per-ctx counter.
__global__ void __launch_bounds__(256) k_p2p_push(P2PDev p, const unsigned long long* __restrict__ src, long long n_u64,
                                                  int parity, unsigned long long seq, int* __restrict__ done) {
  const long long slot = ((long long)parity * p.world + p.rank) * p.slot_stride;
  const long long n2 = n_u64 >> 1;  // KS is even: whole 16-byte vectors
  const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(src);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const ulonglong2 v = s2[i];
    for (int g = 0; g < p.world; ++g) reinterpret_cast<ulonglong2*>(p.peer[g] + slot)[i] = v;
  }
  __threadfence_system();  // this CTA's peer stores are visible system-wide before it reports
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(done, 1);
    if (prev == (int)gridDim.x - 1) {  // last CTA: everybody's data is out
      __threadfence_system();
      for (int g = 0; g < p.world; ++g)
        st_release_sys_u64(p.peer[g] + p.flags_off + ((long long)parity * p.world + p.rank) * P2P_FLAG_STRIDE, seq);
      *done = 0;
    }
  }
}

// One warp; lane r waits for source r.  timeout_cycles bounds the spin (a peer that never arrives must
// not hang this GPU): on expiry *err = 1 and the stream carries on (rbgtopo_fetch reports the error).
__global__ void k_p2p_wait(P2PDev p, int parity, unsigned long long seq, long long timeout_cycles, int* __restrict__ err) {
  const int r = threadIdx.x;
  if (r >= p.world) return;
  const unsigned long long* f = p.peer[p.rank] + p.flags_off + ((long long)parity * p.world + r) * P2P_FLAG_STRIDE;
  const long long t0 = clock64();
  while (ld_acquire_sys_u64(f) < seq) {
    if (clock64() - t0 > timeout_cycles) {
      *err = 1;
      break;
    }
    __nanosleep(64);
  }
}

}  // namespace rbgtopo
