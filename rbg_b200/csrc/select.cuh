// select.cuh — merge of per-chunk / per-rank top-K lists, exclusive-domain
// reselect and the deterministic greedy assignment.
// One warp per step; steps are independent (snapshot semantics, spec §3.7).
#pragma once
#include "kernels.cuh"

namespace rbgtopo {

constexpr int SEL_WARPS = 4;
constexpr int SEL_THREADS = SEL_WARPS * 32;

// Largest key strictly below `prev` among count lists of K keys each
// (lists[i*stride + k]); every lane scans a strided share.
__device__ __forceinline__ unsigned long long next_below(const unsigned long long* base,
                                                         int nlists, long long stride, int K,
                                                         unsigned long long prev) {
  const int lane = threadIdx.x & 31;
  unsigned long long best = 0;
  const int total = nlists * K;
  for (int i = lane; i < total; i += 32) {
    unsigned long long k = base[(long long)(i / K) * stride + (i % K)];
    if (k < prev && k > best) best = k;
  }
  return warp_max_u64(best);
}

// ---- B1: merged[rolerow][KS] = top-K over parts x lc chunk lists; D* ----------
__global__ void __launch_bounds__(SEL_THREADS) k_merge(TopoDev t, BatchDev b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int step = blockIdx.x * SEL_WARPS + warp;
  if (step >= b.n_steps) return;
  const int* hdr = b.blob + RBGTOPO_HDR_WORDS + (size_t)step * RBGTOPO_STEP_WORDS;
  const int flags = hdr[1], fixed_domain = hdr[2], P = hdr[3], role_off = hdr[4];
  const int rolerow_off = hdr[13];
  int kacc = 0;
  int dstar = -1;
  bool dstar_set = false;
  const bool excl_step = (flags & RBGTOPO_STEP_EXCLUSIVE) != 0;
  if (excl_step && fixed_domain >= 0) {
    dstar = fixed_domain;
    dstar_set = true;
  }
  for (int p = 0; p < P; ++p) {
    kacc += b.blob[role_off + 4 * p];
    const int K = min(kacc, t.n);  // spec §3.5: replicas up to and including role p
    unsigned long long* out = b.merged + (size_t)(rolerow_off + p) * KS;
    unsigned long long prev = ~0ull, top = 0;
    int r = 0;
    for (; r < K; ++r) {
      unsigned long long best = 0;
      for (int g = 0; g < b.parts; ++g) {
        const unsigned long long* src =
            b.lists_all + (long long)g * b.part_stride + (size_t)(rolerow_off + p) * b.lc * KS;
        unsigned long long m = next_below(src, b.lc, KS, K, prev);
        best = m > best ? m : best;
      }
      if (best == 0) break;
      if (lane == 0) out[r] = best;
      if (r == 0) top = best;
      prev = best;
    }
    for (int q = r + lane; q < KS; q += 32) out[q] = 0;  // zero padding after the r valid keys
    if (excl_step && !dstar_set && (b.blob[role_off + 4 * p + 3] & RBGTOPO_ROLE_EXCLUSIVE)) {
      // the FIRST participating role decides (spec §3.5)
      dstar = top ? t.domain[key_node(top)] : -1;
      dstar_set = true;
    }
  }
  if (lane == 0) b.dstar[step] = excl_step ? dstar : -1;
}

// ---- B2: restricted reselect inside D* for exclusive steps without a fixed
// domain: scan the nodes of D* that fall into this rank's slab, read the scores
// back from the dense matrix (just written, L2-resident), top-K per role row.
__global__ void __launch_bounds__(SEL_THREADS) k_excl_reselect(TopoDev t, BatchDev b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int step = blockIdx.x * SEL_WARPS + warp;
  if (step >= b.n_steps) return;
  const int* hdr = b.blob + RBGTOPO_HDR_WORDS + (size_t)step * RBGTOPO_STEP_WORDS;
  const int flags = hdr[1], fixed_domain = hdr[2], P = hdr[3], role_off = hdr[4];
  const int rep_off = hdr[12], rolerow_off = hdr[13];
  if (!(flags & RBGTOPO_STEP_EXCLUSIVE) || fixed_domain >= 0) return;
  const int dstar = b.dstar[step];
  int rowbase = 0, kacc = 0;
  for (int p = 0; p < P; ++p) {
    const int count = b.blob[role_off + 4 * p];
    kacc += count;
    const int K = min(kacc, t.n);
    const bool rexcl = (b.blob[role_off + 4 * p + 3] & RBGTOPO_ROLE_EXCLUSIVE) != 0;
    unsigned long long* out = b.excl + (size_t)(rolerow_off + p) * KS;
    if (rexcl) {
      int r = 0;
      if (dstar >= 0) {
        const int d0 = t.dom_ptr[dstar], d1 = t.dom_ptr[dstar + 1];
        const float* row = b.matrix + (size_t)(rep_off + rowbase) * t.slab_stride;
        unsigned long long prev = ~0ull;
        for (; r < K; ++r) {
          unsigned long long best = 0;
          for (int i = d0 + lane; i < d1; i += 32) {
            const int n = t.dom_nodes[i];
            if (n >= t.slab_lo && n < t.slab_hi) {
              const float x = row[n - t.slab_lo];
              if (x != -INFINITY) {
                unsigned long long k = make_key(x, n);
                if (k < prev && k > best) best = k;
              }
            }
          }
          best = warp_max_u64(best);
          if (best == 0) break;
          if (lane == 0) out[r] = best;
          prev = best;
        }
      }
      for (int q = r + lane; q < KS; q += 32) out[q] = 0;
    }
    rowbase += count;
  }
}

// ---- B3: final lists + greedy in replica order (spec §3.6) --------------------
__global__ void __launch_bounds__(SEL_THREADS) k_greedy(TopoDev t, BatchDev b) {
  __shared__ unsigned long long sList[SEL_WARPS][MAXP][KS];
  __shared__ int sTakenNode[SEL_WARPS][KS];
  __shared__ int sTakenAmt[SEL_WARPS][KS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int step = blockIdx.x * SEL_WARPS + warp;
  if (step >= b.n_steps) return;
  const int* hdr = b.blob + RBGTOPO_HDR_WORDS + (size_t)step * RBGTOPO_STEP_WORDS;
  const int flags = hdr[1], fixed_domain = hdr[2], P = hdr[3], role_off = hdr[4];
  const int n_cons = hdr[9], cons_off = hdr[10];
  const int R = hdr[11], rep_off = hdr[12], rolerow_off = hdr[13];
  const bool excl_unknown = (flags & RBGTOPO_STEP_EXCLUSIVE) && fixed_domain < 0;
  const int* con = b.blob + cons_off;

  // final list per role row
  int kacc0 = 0;
  for (int p = 0; p < P; ++p) {
    kacc0 += b.blob[role_off + 4 * p];
    const int K = min(kacc0, t.n);
    const bool rexcl = (b.blob[role_off + 4 * p + 3] & RBGTOPO_ROLE_EXCLUSIVE) != 0;
    if (excl_unknown && rexcl) {
      // merge the per-rank restricted lists
      unsigned long long prev = ~0ull;
      int r = 0;
      for (; r < K; ++r) {
        unsigned long long best = 0;
        for (int g = lane; g < b.parts * K; g += 32) {
          unsigned long long k =
              b.excl_all[(long long)(g / K) * b.excl_part_stride + (size_t)(rolerow_off + p) * KS + (g % K)];
          if (k < prev && k > best) best = k;
        }
        best = warp_max_u64(best);
        if (best == 0) break;
        if (lane == 0) sList[warp][p][r] = best;
        prev = best;
      }
      for (int q = r + lane; q < KS; q += 32) sList[warp][p][q] = 0;
      // publish the final list for inspection (rbgtopo_read_topk)
      __syncwarp();
      b.merged[(size_t)(rolerow_off + p) * KS + lane] = sList[warp][p][lane];
    } else {
      sList[warp][p][lane] = b.merged[(size_t)(rolerow_off + p) * KS + lane];
    }
  }
  __syncwarp();

  int ntaken = 0, unplaced = 0, r = 0;
  for (int p = 0; p < P; ++p) {
    const int count = b.blob[role_off + 4 * p], demand = b.blob[role_off + 4 * p + 1];
    for (int c = 0; c < count; ++c, ++r) {
      int pick = -1;
      for (int k = 0; k < KS; ++k) {  // the list holds K_p keys, then zeros
        const unsigned long long key = sList[warp][p][k];
        if (key == 0) break;
        const int node = key_node(key);
        int used = 0;
        for (int i = lane; i < n_cons; i += 32)
          if (con[2 * i] == node) used += con[2 * i + 1];
        for (int i = lane; i < ntaken; i += 32)
          if (sTakenNode[warp][i] == node) used += sTakenAmt[warp][i];
        used = __reduce_add_sync(FULL, used);
        if (t.free_[node] - used >= demand) {
          pick = node;
          break;
        }
      }
      if (pick >= 0) {
        if (lane == 0) {
          sTakenNode[warp][ntaken] = pick;
          sTakenAmt[warp][ntaken] = demand;
        }
        ++ntaken;
        __syncwarp();
      } else {
        ++unplaced;
      }
      if (lane == 0) b.assign[rep_off + r] = pick;
    }
  }
  __syncwarp();
  int status = unplaced ? RBGTOPO_PLACED_PART : RBGTOPO_PLACED_ALL;
  if (unplaced && (flags & RBGTOPO_STEP_GANG)) {
    status = RBGTOPO_GANG_FAILED;
    for (int i = lane; i < R; i += 32) b.assign[rep_off + i] = -1;
  }
  if (lane == 0) {
    b.status[step] = status;
    b.domain_out[step] = b.dstar[step];
  }
}

}  // namespace rbgtopo
