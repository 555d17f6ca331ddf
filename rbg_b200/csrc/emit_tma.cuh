// emit_tma.cuh — k_emit_tma: the dense (replica x node) matrix of a multi-wave plan written with
// TMA bulk stores from shared memory (cp.async.bulk.global.shared::cta), DESIGN.md §4.2.
//
// Why a second emit kernel.  k_score_emit (score.cuh) streams the matrix with per-thread
// st.global.cs.v4: it needs 6 CTAs x 256 threads per SM (61 K registers) and 62 % of the issue slots
// to keep HBM busy, so nothing else fits on the SM beside it — running the selection kernel
// concurrently made the step SLOWER (99 us vs 91 us serial, profiles/README.md round 2).  Here one
// elected lane issues a 2 KB bulk store per replica row and the TMA engine moves the bytes: 8 warps
// per SM (one 256-thread CTA, 34 KB of shared memory) reach the same write bandwidth
// (profiles/microbench/tma_fill.cu: 6.3 TB/s at 8 warps/SM vs 6.7 TB/s for the plain-store fill),
// and 3/4 of the SM's registers and shared memory are left for k_plan_group, which runs BESIDE this
// kernel on a second stream.
//
// Structure: persistent grid (one CTA per SM), every WARP is an independent worker.
//   item  = a sub-chunk of EMIT_SUB = 512 nodes of this rank's slab x a block of `bsteps` consecutive
//           steps; items are taken from a global atomic counter (dynamic balance: a static split was
//           18 % slower in round 1, the slowest SM sets the time), the index of the next item is
//           fetched while the current one is processed;
//   setup = the sub-chunk's node operands (base, free: 16 floats / ints per lane) into registers and
//           the steps' emit records (emit table: 12 words per step, written once when the batch is
//           staged) into the warp's shared memory — one round of loads per item;
//   tile  = one role row of one step over the sub-chunk: S = need * base where the node is feasible
//           (free >= demand, exclusive roles: domain unowned or ours), else -inf, computed from
//           registers into the warp's private ring of EMIT_STAGES x 2 KB, fence.proxy.async,
//           __syncwarp, then lane 0 issues one bulk store per replica of the role and commits the
//           group; cp.async.bulk.wait_group.read frees a stage for reuse.  No block barrier anywhere.
// Bit-identical to k_score_emit<false> (same expression per element; tests/test_gpu_*).
#pragma once
#include "kernels.cuh"

namespace rbgtopo {

constexpr int EMIT_SUB = 512;                 // nodes per item (2 KB per row store)
constexpr int EMIT_STAGES = 2;                // ring depth per warp
constexpr int EMIT_WARPS = 8;                 // warps per CTA (one CTA per SM)
constexpr int EMIT_TAB_WORDS = 12;            // emit record per step: gid, flags, P, rep_off, 8 packed roles
constexpr int EMIT_MAX_BSTEPS = 8;

__host__ __device__ inline size_t emit_tma_smem_bytes() {
  return (size_t)EMIT_WARPS * (EMIT_STAGES * EMIT_SUB * 4 + EMIT_MAX_BSTEPS * EMIT_TAB_WORDS * 4);
}

// packed role: count | need << 6 | exclusive << 11 | demand << 12   (count <= 32, need <= 16, demand <= 32767)
__device__ __forceinline__ int emit_pack_role(int count, int demand, int need, int flags) {
  return count | (need << 6) | ((flags & RBGTOPO_ROLE_EXCLUSIVE) << 11) | (demand << 12);
}

// One thread per step: blob -> emit table.  Runs once per staged batch (after the blob / the expanded
// plan is in HBM); the table depends on nothing the waves change.
__global__ void k_emit_table(const int* __restrict__ blob, int n_steps, int* __restrict__ etab) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_steps) return;
  const int* hdr = blob + RBGTOPO_HDR_WORDS + (size_t)s * RBGTOPO_STEP_WORDS;
  const int P = hdr[3];
  const int* roles = blob + hdr[4];
  int* e = etab + (size_t)s * EMIT_TAB_WORDS;
  e[0] = hdr[0];
  e[1] = hdr[1];
  e[2] = P;
  e[3] = hdr[12];
  for (int p = 0; p < MAXP; ++p)
    e[4 + p] = p < P ? emit_pack_role(roles[4 * p], roles[4 * p + 1], roles[4 * p + 2], roles[4 * p + 3]) : 0;
}

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ctr[0] = item queue, ctr[1] = warps that left the loop (the last one resets both for the next launch)
__global__ void __launch_bounds__(32 * EMIT_WARPS, 4)
k_emit_tma(TopoDev t, BatchDev b, const int* __restrict__ etab, int subs, int items, int* __restrict__ ctr) {
  extern __shared__ __align__(128) unsigned char em_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* const ring = reinterpret_cast<float*>(em_smem) + (size_t)warp * EMIT_STAGES * EMIT_SUB;
  int* const sTab = reinterpret_cast<int*>(em_smem + (size_t)EMIT_WARPS * EMIT_STAGES * EMIT_SUB * 4) +
                    warp * EMIT_MAX_BSTEPS * EMIT_TAB_WORDS;
  const int BS = b.bsteps;
  const size_t stride = (size_t)t.slab_stride;
  constexpr int V = EMIT_SUB / 128;  // float4 groups per lane

  int it = 0;
  int next = 0;
  if (lane == 0) next = atomicAdd(&ctr[0], 1);
  next = __shfl_sync(FULL, next, 0);
  while (next < items) {
    const int item = next;
    if (lane == 0) next = atomicAdd(&ctr[0], 1);  // in flight while this item is processed
    const int blk = item / subs, sub = item - blk * subs;
    const int step0 = blk * BS;
    const int nst = min(b.n_steps, step0 + BS) - step0;
    const int n0 = t.slab_lo + sub * EMIT_SUB;
    const int n1 = min(n0 + EMIT_SUB, t.slab_hi);
    const int len4 = (n1 - n0 + 3) >> 2;  // float4 groups of the sub-chunk (the last one is padded with -inf)
    // ---- setup: emit records -> shared memory, node operands -> registers (one round of loads)
    __syncwarp();  // the previous item's readers of sTab are done
    for (int i = lane; i < nst * (EMIT_TAB_WORDS / 4); i += 32)
      reinterpret_cast<int4*>(sTab)[i] = __ldg(reinterpret_cast<const int4*>(etab + (size_t)step0 * EMIT_TAB_WORDS) + i);
    float4 base4[V];
    int4 av[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int g = lane + 32 * j;
      base4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      av[j] = make_int4(-1, -1, -1, -1);
      if (g < len4) {
        const int n = n0 + (g << 2);
        base4[j] = __ldg(reinterpret_cast<const float4*>(t.base + n));
        av[j] = __ldg(reinterpret_cast<const int4*>(t.free_ + n));  // padded past n: safe
        if (n + 4 > n1) {  // only in the slab's last group: lanes past the slab are infeasible
          if (n + 1 >= n1) av[j].y = -1;
          if (n + 2 >= n1) av[j].z = -1;
          av[j].w = -1;
        }
      }
    }
    __syncwarp();
    float* const col0 = b.matrix + (n0 - t.slab_lo);
    const uint32_t row_bytes = (uint32_t)len4 * 16u;
    for (int si = 0; si < nst; ++si) {
      const int* e = sTab + si * EMIT_TAB_WORDS;
      const int gid = e[0], P = e[2];
      const bool excl_step = (e[1] & RBGTOPO_STEP_EXCLUSIVE) != 0;
      float* rowp = col0 + (size_t)e[3] * stride;
      for (int p = 0; p < P; ++p) {
        const int pr = e[4 + p];
        const int count = pr & 63, demand = pr >> 12;
        const float need = (float)((pr >> 6) & 31);
        const bool rexcl = (pr >> 11) & 1;
        float* const st = ring + (size_t)(it % EMIT_STAGES) * EMIT_SUB;
        if (lane == 0) bulk_wait_read<EMIT_STAGES - 1>();  // the stores that read this stage are done with it
        __syncwarp();
#pragma unroll
        for (int j = 0; j < V; ++j) {
          int4 a = av[j];
          if (rexcl && excl_step && lane + 32 * j < len4) {  // exclusive roles: domains owned by another group are infeasible
            const int4 ow = __ldg(reinterpret_cast<const int4*>(t.node_owner + n0 + ((lane + 32 * j) << 2)));
            if (!(ow.x == -1 || ow.x == gid)) a.x = -1;
            if (!(ow.y == -1 || ow.y == gid)) a.y = -1;
            if (!(ow.z == -1 || ow.z == gid)) a.z = -1;
            if (!(ow.w == -1 || ow.w == gid)) a.w = -1;
          }
          float4 o4;
          o4.x = a.x >= demand ? need * base4[j].x : -INFINITY;
          o4.y = a.y >= demand ? need * base4[j].y : -INFINITY;
          o4.z = a.z >= demand ? need * base4[j].z : -INFINITY;
          o4.w = a.w >= demand ? need * base4[j].w : -INFINITY;
          reinterpret_cast<float4*>(st)[lane + 32 * j] = o4;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          for (int c = 0; c < count; ++c) bulk_s2g(rowp + (size_t)c * stride, st, row_bytes);
          bulk_commit();
        }
        rowp += (size_t)count * stride;
        ++it;
      }
    }
    next = __shfl_sync(FULL, next, 0);
  }
  if (lane == 0) {
    bulk_wait_all();  // every store of this warp has been written
    const int total = (int)(gridDim.x * EMIT_WARPS);
    if (atomicAdd(&ctr[1], 1) == total - 1) {  // everybody left the loop: re-arm the queue
      ctr[0] = 0;
      ctr[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace rbgtopo
