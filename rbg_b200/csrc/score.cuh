// score.cuh — k_score_emit: the dominant kernel of the path.
//
// Emits the dense (replica x node) score matrix as a pure HBM write stream.
// One work item = (step, chunk of `chunk` <= 2048 nodes of this rank's slab); a
// persistent grid of 256-thread CTAs takes contiguous item ranges, so the chunks
// of a step run back to back on one SM (its header, roles, anchors and the
// anchors' CSR rows are L1 hits after the first chunk).  No shared memory.
//   1. background: every role row is  S = need*base[n]  where the node is feasible
//      (free >= demand, and for exclusive roles the domain is unowned or ours),
//      else -inf; written once per replica of the role with 128-bit streaming
//      stores.  base/free are per-snapshot vectors shared by all steps (L1/L2).
//   2. only for steps with anchor pods / consumed capacity, after one block
//      barrier: a warp walks the CSR row of anchor m (coalesced int32 loads) and
//      adds pair*c*w onto the just-written, L2-hot scores of the neighbours inside
//      the chunk with fire-and-forget red.global.add.f32 (+ the self term); nodes
//      whose consumed capacity makes them infeasible are overwritten with -inf.
//      All addends are exact integers and -inf absorbs adds, so the result is
//      bit-identical to the oracle's sequential fp32 accumulation in any order.
// Selection never touches this kernel's data path except for reading back the
// few patched scores (select.cuh).
#pragma once
#include "kernels.cuh"

namespace rbgtopo {

constexpr int GPT = 2;  // float4 groups per thread: chunk <= 256 * 4 * GPT = 2048

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(SCORE_THREADS, 6)
k_score_emit(TopoDev t, BatchDev b, int items) {
  const int T = b.chunk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int groups = T >> 2;
  const int* __restrict__ blob = b.blob;
  const size_t stride = (size_t)t.slab_stride;

  const int per = (items + gridDim.x - 1) / gridDim.x;
  int item = blockIdx.x * per;
  const int item_end = min(items, item + per);
  if (item >= item_end) return;
  int step = item / b.lc, ch = item - step * b.lc;
  for (; item < item_end; ++item) {
    const int n0 = t.slab_lo + ch * T;
    const int n1 = min(n0 + T, t.slab_hi);
    const int* __restrict__ hdr = blob + RBGTOPO_HDR_WORDS + (size_t)step * RBGTOPO_STEP_WORDS;
    const int4 h0 = __ldg(reinterpret_cast<const int4*>(hdr));      // gid flags fixed P
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(hdr) + 1);  // role_off Q pair_off n_anchors
    const int4 h2 = __ldg(reinterpret_cast<const int4*>(hdr) + 2);  // anchor_off n_cons cons_off R
    const int4 h3 = __ldg(reinterpret_cast<const int4*>(hdr) + 3);  // rep_off rolerow_off - -
    const int gid = h0.x, P = h0.w;
    const bool excl_step = (h0.y & RBGTOPO_STEP_EXCLUSIVE) != 0;
    const int4* __restrict__ roles = reinterpret_cast<const int4*>(blob + h1.x);  // 16-byte aligned (validated)
    float* const mrow0 = b.matrix + (size_t)h3.x * stride + (n0 - t.slab_lo);

    // ---- 1. background rows
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
      const int g = tid + j * SCORE_THREADS;
      const int n = n0 + (g << 2);
      if (g < groups && n < n1) {
        const float4 base4 = __ldg(reinterpret_cast<const float4*>(t.base + n));
        const int4 av = __ldg(reinterpret_cast<const int4*>(t.free_ + n));  // padded past n: safe
        uint32_t okm = 0xFu;
        if (n + 4 > n1) okm = (1u << (n1 - n)) - 1u;  // only in the slab's last group
        uint32_t blk = 0;
        if (excl_step) {
          const int4 ow = __ldg(reinterpret_cast<const int4*>(t.node_owner + n));
          blk = (!(ow.x == -1 || ow.x == gid) ? 1u : 0u) | (!(ow.y == -1 || ow.y == gid) ? 2u : 0u) |
                (!(ow.z == -1 || ow.z == gid) ? 4u : 0u) | (!(ow.w == -1 || ow.w == gid) ? 8u : 0u);
        }
        float* rowp = mrow0 + (g << 2);
        for (int p = 0; p < P; ++p) {
          const int4 role = __ldg(roles + p);  // count demand need flags
          const float need = (float)role.z;
          const uint32_t bad = (excl_step && (role.w & RBGTOPO_ROLE_EXCLUSIVE)) ? blk : 0u;
          const uint32_t good = okm & ~bad;
          float4 o4;
          o4.x = (av.x >= role.y && (good & 1u)) ? need * base4.x : -INFINITY;
          o4.y = (av.y >= role.y && (good & 2u)) ? need * base4.y : -INFINITY;
          o4.z = (av.z >= role.y && (good & 4u)) ? need * base4.z : -INFINITY;
          o4.w = (av.w >= role.y && (good & 8u)) ? need * base4.w : -INFINITY;
          for (int c = 0; c < role.x; ++c) {
            st_stream_f4(rowp, o4);
            rowp += stride;
          }
        }
      }
    }

    // ---- 2. sparse corrections (steps with anchors / consumed capacity only)
    if ((h1.w | h2.y) != 0) {
      __threadfence();
      __syncthreads();  // the chunk's background scores are written
      const int* __restrict__ con = blob + h2.z;
      for (int c = tid; c < h2.y; c += SCORE_THREADS) {  // consumed capacity -> maybe infeasible
        const int m = __ldg(con + 2 * c);
        if (m >= n0 && m < n1) {
          int amt = 0;
          for (int k = 0; k < h2.y; ++k)
            if (__ldg(con + 2 * k) == m) amt += __ldg(con + 2 * k + 1);  // duplicates add up
          const int avail = __ldg(t.free_ + m) - amt;
          float* rowp = mrow0 + (m - n0);
          for (int p = 0; p < P; ++p) {
            const int4 role = __ldg(roles + p);
            if (avail < role.y)
              for (int k = 0; k < role.x; ++k) rowp[(size_t)k * stride] = -INFINITY;
            rowp += (size_t)role.x * stride;
          }
        }
      }
      const int* __restrict__ anc = blob + h2.x;
      const int Q = h1.y;
      for (int a = warp; a < h1.w; a += SCORE_WARPS) {  // one warp per anchor pod
        const int m = __ldg(anc + 3 * a), q = __ldg(anc + 3 * a + 1), c = __ldg(anc + 3 * a + 2);
        const int rb = __ldg(t.row_ptr + m), re = __ldg(t.row_ptr + m + 1);
        for (int j = rb + lane; j <= re; j += 32) {  // j == re stands for the self term
          int nn, wv;
          if (j < re) {
            nn = __ldg(t.col + j);
            wv = __ldg(t.w + j) * c;
          } else {
            nn = m;
            wv = RBGTOPO_SELF_W * c;
          }
          if (nn >= n0 && nn < n1) {
            float* rowp = mrow0 + (nn - n0);
            for (int p = 0; p < P; ++p) {
              const int count = __ldg(blob + h1.x + 4 * p);
              const int coef = __ldg(blob + h1.z + p * Q + q);
              if (coef) {
                const float add = (float)(coef * wv);
                for (int k = 0; k < count; ++k) red_add_f32(rowp + (size_t)k * stride, add);
              }
              rowp += (size_t)count * stride;
            }
          }
        }
      }
    }
    if (++ch == b.lc) {
      ch = 0;
      ++step;
    }
  }
}

}  // namespace rbgtopo
