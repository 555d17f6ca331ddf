// score.cuh — k_score_select: the dominant kernel of the path.
//
// One work item = (step, chunk of `chunk` <= 2048 nodes of this rank's slab);
// a persistent grid of 256-thread CTAs strides over the items.
//   1. load the step record, zero the per-role delta rows in shared memory, load
//      capacity and (exclusive steps) the domain-ownership mask
//   2. scatter the step's anchor pods into the deltas: one warp walks CSR row m
//      (coalesced int32 loads) and adds pair*c*w to the delta of every neighbour
//      inside the chunk (+ the self term); consumed capacity is subtracted
//   3. per role row: S = need*base + delta (one FFMA per score, exact), mask
//      infeasible -> -inf, write the row once per replica of the role with
//      128-bit streaming stores.  Each thread keeps its <= 8 scores of the role
//      in registers as 32-bit local keys  (int(S) << 3 | 7 - e)  so that
//   4. every warp selects the exact top-K of its 256 scores with K REDUX rounds
//      (warp max of the lane maxima; the winning lane rescans its own registers
//      for its next best) — no shared memory, no block barrier.
//   5. after the last role one barrier; warp p merges the 8 per-warp lists of
//      role p (<= 8*K keys) and writes the chunk's top-K keys to global memory.
// K of role p = number of replicas of the step up to and including role p
// (spec §3.5: earlier replicas can exhaust at most that many - 1 nodes).
#pragma once
#include "kernels.cuh"

namespace rbgtopo {

constexpr int GPT = 2;  // float4 groups per thread: chunk <= 256 * 4 * GPT = 2048
constexpr int EPT = 4 * GPT;

// host mirror: rbgtopo.cu score_smem_bytes()
//   sD[PB][T] f32 | sAvail[T] i32 | sBlk[T/32 (+pad to 8 B)] u32 | sWin[PB][8][KS] u64
__global__ void __launch_bounds__(SCORE_THREADS, 4)
k_score_select(TopoDev t, BatchDev b, int items, int PB /* max roles per step in the batch */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sHdr[RBGTOPO_STEP_WORDS];
  __shared__ int sRole[MAXP * 4];
  __shared__ int sPair[MAXP * MAXQ];
  const int T = b.chunk;
  float* sD = reinterpret_cast<float*>(smem_raw);
  int* sAvail = reinterpret_cast<int*>(sD + (size_t)PB * T);
  uint32_t* sBlk = reinterpret_cast<uint32_t*>(sAvail + T);
  unsigned long long* sWin =
      reinterpret_cast<unsigned long long*>(sBlk + (T >> 5) + ((T >> 5) & 1));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int groups = T >> 2;

  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int step = item / b.lc, ch = item - step * b.lc;
    const int n0 = t.slab_lo + ch * T;
    const int n1 = min(n0 + T, t.slab_hi);
    __syncthreads();  // previous item's shared memory is dead
    if (tid < RBGTOPO_STEP_WORDS)
      sHdr[tid] = b.blob[RBGTOPO_HDR_WORDS + (size_t)step * RBGTOPO_STEP_WORDS + tid];
    __syncthreads();
    const int gid = sHdr[0], flags = sHdr[1], fixed_domain = sHdr[2], P = sHdr[3];
    const int Q = sHdr[5];
    const int n_anchors = sHdr[7], n_cons = sHdr[9];
    const int rep_off = sHdr[12], rolerow_off = sHdr[13];
    const bool excl_step = (flags & RBGTOPO_STEP_EXCLUSIVE) != 0;

    // ---- 1. step parameters + init (independent, one barrier)
    if (tid < P * 4) sRole[tid] = b.blob[sHdr[4] + tid];
    for (int i = tid; i < P * MAXQ; i += SCORE_THREADS) {
      const int p = i / MAXQ, q = i - p * MAXQ;
      sPair[i] = (q < Q) ? b.blob[sHdr[6] + p * Q + q] : 0;
    }
    {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = tid; i < P * groups; i += SCORE_THREADS) reinterpret_cast<float4*>(sD)[i] = z;
      // P * groups float4 == the first P rows of sD (rows are T floats = groups float4)
    }
    for (int i = tid; i < T; i += SCORE_THREADS) {
      const int n = n0 + i;
      int av = -1;
      bool blk = false;
      if (n < n1) {
        av = t.free_[n];
        if (excl_step) {
          const int o = t.node_owner[n];
          blk = !(o == -1 || o == gid);
        }
      }
      sAvail[i] = av;
      const uint32_t bm = __ballot_sync(FULL, blk);
      if (lane == 0) sBlk[i >> 5] = bm;
    }
    __syncthreads();

    // ---- 2. anchors (one warp per anchor pod) and consumed capacity
    {
      const int* anc = b.blob + sHdr[8];
      for (int a = warp; a < n_anchors; a += SCORE_WARPS) {
        const int m = anc[3 * a], q = anc[3 * a + 1], c = anc[3 * a + 2];
        const int rb = t.row_ptr[m], re = t.row_ptr[m + 1];
        for (int j = rb + lane; j < re; j += 32) {
          const int nn = t.col[j];
          if (nn >= n0 && nn < n1) {
            const int wv = t.w[j] * c;
            for (int p = 0; p < P; ++p) {
              const int coef = sPair[p * MAXQ + q];
              if (coef) atomicAdd(&sD[p * T + (nn - n0)], (float)(coef * wv));
            }
          }
        }
        if (lane == 0 && m >= n0 && m < n1) {
          for (int p = 0; p < P; ++p) {
            const int coef = sPair[p * MAXQ + q] * c;
            if (coef) atomicAdd(&sD[p * T + (m - n0)], (float)(coef * RBGTOPO_SELF_W));
          }
        }
      }
      const int* con = b.blob + sHdr[10];
      for (int c = tid; c < n_cons; c += SCORE_THREADS) {
        const int m = con[2 * c];
        if (m >= n0 && m < n1) atomicSub(&sAvail[m - n0], con[2 * c + 1]);
      }
    }
    __syncthreads();

    // ---- 3+4. per role: stream the row, select per warp
    const bool write_rows = b.emit_matrix || (excl_step && fixed_domain < 0);
    const bool restrict_fixed = excl_step && fixed_domain >= 0;
    // thread-invariant per-group data, loaded once for all roles
    float4 base4[GPT];
    int4 av4[GPT];
    uint32_t okbits[GPT];  // bit i: lane valid; bit 4+i: blocked by ownership; bit 8+i: in fixed domain
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
      const int g = tid + j * SCORE_THREADS;
      const int n = n0 + (g << 2);
      okbits[j] = 0;
      base4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      av4[j] = make_int4(-1, -1, -1, -1);
      if (g < groups && n < n1) {
        base4[j] = *reinterpret_cast<const float4*>(t.base + n);
        av4[j] = *reinterpret_cast<const int4*>(sAvail + (g << 2));
        const uint32_t blk = (sBlk[g >> 3] >> ((g & 7) << 2)) & 0xFu;
        uint32_t valid = 0, dom = 0xFu;
#pragma unroll
        for (int i = 0; i < 4; ++i) valid |= (n + i < n1) ? (1u << i) : 0u;
        if (restrict_fixed) {
          dom = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n + i < n1 && t.domain[n + i] == fixed_domain) dom |= 1u << i;
        }
        okbits[j] = valid | (blk << 4) | (dom << 8);
      }
    }

    int kacc = 0, rowbase = 0;
    for (int p = 0; p < P; ++p) {
      const int count = sRole[4 * p], demand = sRole[4 * p + 1];
      const float need = (float)sRole[4 * p + 2];
      const bool rexcl = excl_step && (sRole[4 * p + 3] & RBGTOPO_ROLE_EXCLUSIVE);
      kacc += count;
      const int Kp = min(kacc, t.n);
      int k32[EPT];
#pragma unroll
      for (int j = 0; j < GPT; ++j) {
        const int g = tid + j * SCORE_THREADS;
        const uint32_t ok = okbits[j];
        float v[4];
        if (ok & 0xFu) {
          const float4 d4 = *reinterpret_cast<const float4*>(sD + p * T + (g << 2));
          const float bb[4] = {base4[j].x, base4[j].y, base4[j].z, base4[j].w};
          const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
          const int av[4] = {av4[j].x, av4[j].y, av4[j].z, av4[j].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = fmaf(need, bb[i], dd[i]);
            const bool feas = (av[i] >= demand) && ((ok >> i) & 1u) && !(rexcl && ((ok >> (4 + i)) & 1u));
            v[i] = feas ? x : -INFINITY;
            const bool sel = feas && !(rexcl && !((ok >> (8 + i)) & 1u));
            k32[j * 4 + i] = sel ? ((__float2int_rn(x) << 3) | (7 - (j * 4 + i))) : -1;
          }
          if (write_rows) {
            const float4 o4 = make_float4(v[0], v[1], v[2], v[3]);
            float* rowp = b.matrix + (size_t)(rep_off + rowbase) * t.slab_stride + (n0 - t.slab_lo) + (g << 2);
            for (int c = 0; c < count; ++c) st_stream_f4(rowp + (size_t)c * t.slab_stride, o4);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) k32[j * 4 + i] = -1;
        }
      }
      rowbase += count;

      // ---- per-warp exact top-Kp from registers
      int cur = k32[0];
#pragma unroll
      for (int e = 1; e < EPT; ++e) cur = max(cur, k32[e]);
      unsigned long long* win = sWin + ((size_t)p * SCORE_WARPS + warp) * KS;
      int r = 0;
      for (; r < Kp; ++r) {
        unsigned long long key = 0;
        if (cur >= 0) {
          const int e = 7 - (cur & 7);
          const int node = n0 + ((tid + (e >> 2) * SCORE_THREADS) << 2) + (e & 3);
          key = make_key((float)(cur >> 3), node);
        }
        const unsigned long long m = warp_max_u64(key);
        if (m == 0) break;
        if (lane == 0) win[r] = m;
        if (key == m) {  // this lane won: its next best is its largest local key below cur
          int nxt = -1;
#pragma unroll
          for (int e = 0; e < EPT; ++e) nxt = (k32[e] < cur) ? max(nxt, k32[e]) : nxt;
          cur = nxt;
        }
      }
      for (int q = r + lane; q < Kp; q += 32) win[q] = 0;
    }
    __syncthreads();

    // ---- 5. warp p merges the 8 per-warp lists of role p
    if (warp < P) {
      const int p = warp;
      int kp = 0;
      for (int q = 0; q <= p; ++q) kp += sRole[4 * q];
      kp = min(kp, t.n);
      const unsigned long long* src = sWin + (size_t)p * SCORE_WARPS * KS;
      unsigned long long* out = b.lists + ((size_t)(rolerow_off + p) * b.lc + ch) * KS;
      unsigned long long prev = ~0ull;
      int r = 0;
      for (; r < kp; ++r) {
        unsigned long long best = 0;
        for (int i = lane; i < SCORE_WARPS * kp; i += 32) {
          const unsigned long long k = src[(i / kp) * KS + (i % kp)];
          if (k < prev && k > best) best = k;
        }
        best = warp_max_u64(best);
        if (best == 0) break;
        if (lane == 0) out[r] = best;
        prev = best;
      }
      for (int q = r + lane; q < KS; q += 32) out[q] = 0;
    }
  }
}

}  // namespace rbgtopo
