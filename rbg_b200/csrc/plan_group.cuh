// plan_group.cuh — k_plan_group: every wave of a group in ONE launch (world == 1
// multi-wave plans).
//
// Waves of one group depend only on each other (batch = snapshot semantics, spec
// §3.7), so one CTA walks its group's steps through the `next_step` links and
// carries the group's state in shared memory instead of chaining it through HBM
// and three launches:
//   * the hash table of patched nodes is INCREMENTAL: wave w only inserts the
//     closed neighbourhoods of the replicas wave w-1 placed.  A slot keeps, per
//     GROUP role q, aw[q] = sum of c*w over the anchor pods of role q next to the
//     node, and the capacity consumed on it; the score delta of a role row is
//     sum_q pair[p][q] * aw[q] (exact integers, spec §3.4), evaluated where needed;
//   * placements, the exclusive domain and a failed gang never leave the CTA;
//   * the dense matrix gets its sparse corrections per wave from the table: one
//     red.global.add.f32 per (patched node, replica row) with the summed delta,
//     -inf where consumed capacity made the node infeasible.
// Selection (patched slots merged with the walk of the background order) and the
// greedy are the ones of select_fast.cuh.  Steps are read from the expanded plan
// blob; the chained anchor / consumed records of later steps are neither written
// nor read here.
#pragma once
#include "select_fast.cuh"

namespace rbgtopo {

struct GroupTab {
  int* node;        // [HT] key, -1 = empty
  int* cons;        // [HT] capacity consumed on the slot's node by earlier waves
  float* aw;        // [QB][HT] anchor weight per group role
  int mask, HT;
  int* cnt;         // occupied slots; the thread whose CAS claims a slot appends it to dSlot
  int* dSlot;       // [CAP] dense view of the occupied slots (insertion order)
  float* dBase;     // [CAP]
  int* dFree;       // [CAP] free capacity of the node in the snapshot
  int* dAvail;      // [CAP] dFree - cons, refreshed per wave
  int* dDom;        // [CAP] domain, bit 31 set = domain owned by another group
};
struct GroupRole {  // role row of the current wave, staged in shared memory
  int count, demand, need, flags;
};

__host__ __device__ inline size_t group_smem_bytes(int QB, int PB, int HT, int CAP) {
  return (size_t)HT * 4 * (2 + QB) + (size_t)CAP * 20 +                          // table + dense view
         (size_t)PB * KS * (3 * 8 + 3 * 4) +                                      // per-role key lists + capacities
         (size_t)MAXP * 16 + (size_t)MAXP * RBGTOPO_MAX_GROUP_ROLES * 4;          // staged roles + pair rows
}

__device__ __forceinline__ int gtab_insert(const GroupTab& T, int n) {
  int h = tab_hash(n, T.mask);
  while (true) {
    const int old = atomicCAS(&T.node[h], -1, n);
    if (old == -1) T.dSlot[atomicAdd(T.cnt, 1)] = h;  // claimed: the capacity bound keeps this below CAP
    if (old == -1 || old == n) return h;
    h = (h + 1) & T.mask;
  }
}
__device__ __forceinline__ bool gtab_has(const GroupTab& T, int n) {
  int h = tab_hash(n, T.mask);
  while (true) {
    const int k = T.node[h];
    if (k == n) return true;
    if (k == -1) return false;
    h = (h + 1) & T.mask;
  }
}
__device__ __forceinline__ float gtab_delta(const GroupTab& T, const float* pair_row, int Q, int slot) {
  float d = 0.0f;
  const float* aw = T.aw + slot;
  for (int q = 0; q < Q; ++q, aw += T.HT) d = fmaf(pair_row[q], *aw, d);
  return d;
}

// top-K of a role row into out[0..KS) (+ capacities): select_role_fast with the
// delta evaluated from the per-group-role planes.  One warp.
__device__ __forceinline__ void select_role_group(const TopoDev& t, int gid, bool excl_step, const GroupRole& role,
                                                  const float* pair_row, int Q, int K, int dom, const GroupTab& T,
                                                  int cnt, unsigned long long* sAcc, int* sAccAv,
                                                  unsigned long long* sPat, int* sPatAv, unsigned long long* out,
                                                  int* outAvail) {
  const int lane = threadIdx.x & 31;
  if (dom == DOM_NONE || K <= 0) {
    out[lane] = 0;
    outAvail[lane] = 0;
    __syncwarp();
    return;
  }
  const int demand = role.demand;
  const float need = (float)role.need;
  const bool rexcl = excl_step && (role.flags & RBGTOPO_ROLE_EXCLUSIVE);

  // ---- (a) patched slots, K strictly-descending rounds.  The keys of the first 32 * EREG
  //      entries are evaluated once and stay in registers; entries past that are re-evaluated
  //      per round.
  constexpr int EREG = 8;
  auto entry_key = [&](int i) -> unsigned long long {
    const int av = T.dAvail[i], dd = T.dDom[i];
    if (av >= demand && !(rexcl && dd < 0) && (dom == DOM_ANY || (dd & 0x7FFFFFFF) == dom)) {
      const int slot = T.dSlot[i];
      return make_key(fmaf(need, T.dBase[i], gtab_delta(T, pair_row, Q, slot)), T.node[slot]);
    }
    return 0ull;
  };
  unsigned long long kreg[EREG];
#pragma unroll
  for (int j = 0; j < EREG; ++j) {
    const int i = lane + 32 * j;
    kreg[j] = i < cnt ? entry_key(i) : 0ull;
  }
  int npat = 0;
  {
    unsigned long long prev = ~0ull;
    for (; npat < K; ++npat) {
      unsigned long long best = 0;
      int bi = 0;
#pragma unroll
      for (int j = 0; j < EREG; ++j)
        if (kreg[j] < prev && kreg[j] > best) { best = kreg[j]; bi = lane + 32 * j; }
      for (int i = lane + 32 * EREG; i < cnt; i += 32) {
        const unsigned long long k = entry_key(i);
        if (k < prev && k > best) { best = k; bi = i; }
      }
      const unsigned long long m = warp_max_u64(best);
      if (m == 0) break;
      if (best == m) { sPat[npat] = m; sPatAv[npat] = T.dAvail[bi]; }  // keys are unique: one lane
      prev = m;
    }
  }
  __syncwarp();

  // ---- (b) walk the background order; patched nodes are skipped by a table probe
  const int slab_len = t.slab_hi - t.slab_lo;
  int acc = 0;
  for (int pos = 0; pos < slab_len && acc < K; pos += 32) {
    const int i = pos + lane;
    int av = 0;
    unsigned long long key = 0;
    bool ok = false;
    if (i < slab_len) {
      int node;
      if (role.need > 0) {
        const unsigned long long ob = t.order[i];
        node = key_node(ob);
        const float base = __uint_as_float((uint32_t)(ob >> 32) ^ 0x80000000u);  // base >= 0
        key = make_key(need * base, node);
      } else {
        node = t.slab_lo + i;
        key = make_key(0.0f, node);
      }
      av = t.free_[node];
      ok = av >= demand;
      if (ok && rexcl) {
        const int o = t.node_owner[node];
        ok = (o == -1 || o == gid);
      }
      if (ok && dom != DOM_ANY) ok = t.domain[node] == dom;
      if (ok) ok = !gtab_has(T, node);
    }
    const uint32_t m = __ballot_sync(FULL, ok);
    const int idx = acc + __popc(m & ((1u << lane) - 1u));
    if (ok && idx < K) { sAcc[idx] = key; sAccAv[idx] = av; }
    acc += __popc(m);
  }
  acc = min(acc, K);
  __syncwarp();

  merge_lists(sAcc, sAccAv, acc, sPat, sPatAv, npat, K, out, outAvail);
}

// Inserts the closed neighbourhood of anchor pod(s) (node m, group role q, count c) and the
// capacity `dem` consumed on m.  One warp.
__device__ __forceinline__ void gtab_add_anchor(const TopoDev& t, const GroupTab& T, int m, int q, int c, int dem) {
  const int lane = threadIdx.x & 31;
  if (lane == 0 && dem > 0 && m >= t.slab_lo && m < t.slab_hi) atomicAdd(&T.cons[gtab_insert(T, m)], dem);
  if (c <= 0) return;
  const int rb = t.row_ptr[m], re = t.row_ptr[m + 1];
  for (int j = rb + lane; j <= re; j += 32) {  // j == re stands for the anchor's own node
    int nn, wv;
    if (j < re) {
      nn = t.col[j];
      wv = t.w[j] * c;
    } else {
      nn = m;
      wv = RBGTOPO_SELF_W * c;
    }
    if (nn >= t.slab_lo && nn < t.slab_hi) atomicAdd(&T.aw[(size_t)q * T.HT + gtab_insert(T, nn)], (float)wv);
  }
}

// grid = groups with at least one pending replica = the steps of wave 0; CTA g starts at step g.
// QB = largest role count of a group in the batch, PB = warps per CTA (>= roles of any wave).
__global__ void __launch_bounds__(32 * MAXP, 4) k_plan_group(TopoDev t, BatchDev b, int QB, int HT, int CAP) {
  extern __shared__ __align__(16) unsigned char pg_smem[];
  __shared__ int sTakenNode[KS], sTakenAmt[KS], sTakenRole[KS];
  __shared__ int sDstar, sCnt, sNew, sStatus, sAny;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int PB = nwarps;
  GroupTab T;
  T.node = reinterpret_cast<int*>(pg_smem);
  T.cons = T.node + HT;
  T.aw = reinterpret_cast<float*>(T.cons + HT);
  T.mask = HT - 1;
  T.HT = HT;
  T.cnt = &sCnt;
  T.dSlot = reinterpret_cast<int*>(T.aw + (size_t)QB * HT);
  T.dBase = reinterpret_cast<float*>(T.dSlot + CAP);
  T.dFree = reinterpret_cast<int*>(T.dBase + CAP);
  T.dAvail = T.dFree + CAP;
  T.dDom = T.dAvail + CAP;
  // per-role lists (8-byte aligned: everything before is a multiple of 8 bytes when CAP, HT are multiples of 32)
  unsigned long long* sList = reinterpret_cast<unsigned long long*>(T.dDom + CAP);
  unsigned long long* sAcc = sList + (size_t)PB * KS;
  unsigned long long* sPat = sAcc + (size_t)PB * KS;
  int* sListAv = reinterpret_cast<int*>(sPat + (size_t)PB * KS);
  int* sAccAv = sListAv + (size_t)PB * KS;
  int* sPatAv = sAccAv + (size_t)PB * KS;
  GroupRole* sRole = reinterpret_cast<GroupRole*>(sPatAv + (size_t)PB * KS);
  float* sPair = reinterpret_cast<float*>(sRole + MAXP);  // [MAXP][RBGTOPO_MAX_GROUP_ROLES]

  for (int i = tid; i < HT; i += nthreads) {
    T.node[i] = -1;
    T.cons[i] = 0;
  }
  for (int i = tid; i < QB * HT; i += nthreads) T.aw[i] = 0.0f;
  if (tid == 0) sCnt = 0;

  int step = blockIdx.x;
  StepHdr h = load_hdr(b, step);
  const bool excl_step = (h.flags & RBGTOPO_STEP_EXCLUSIVE) != 0;
  const bool gang = (h.flags & RBGTOPO_STEP_GANG) != 0;
  const int gid = h.gid, Q = h.Q;
  int fixed = excl_step ? h.fixed_domain : -1;
  const size_t stride = (size_t)t.slab_stride;
  __syncthreads();

  // the group's scheduled pods (anchor records of its first step)
  {
    const int* anc = b.blob + h.anchor_off;
    for (int a = warp; a < h.n_anchors - h.i0; a += nwarps) gtab_add_anchor(t, T, anc[3 * a], anc[3 * a + 1], anc[3 * a + 2], 0);
  }
  int n_new = 0;     // replicas placed by the previous wave: sTaken*[0, n_new)
  int cnt_done = 0;  // dense entries whose node attributes are loaded

  while (true) {
    // ---- A. closed neighbourhoods + consumption of the previous wave's placements; this wave's roles
    for (int a = warp; a < n_new; a += nwarps) gtab_add_anchor(t, T, sTakenNode[a], sTakenRole[a], 1, sTakenAmt[a]);
    if (tid < h.P) {
      const int4 r = *reinterpret_cast<const int4*>(b.blob + h.role_off + 4 * tid);
      sRole[tid] = GroupRole{r.x, r.y, r.z, r.w};
    }
    for (int i = tid; i < h.P * Q; i += nthreads)
      sPair[(i / Q) * RBGTOPO_MAX_GROUP_ROLES + i % Q] = (float)b.blob[h.pair_off + i];
    __syncthreads();

    // ---- B. node attributes of the slots claimed since the last wave, capacities of all
    const int cnt = sCnt;
    for (int d = cnt_done + tid; d < cnt; d += nthreads) {
      const int node = T.node[T.dSlot[d]];
      int dd = t.domain[node];
      if (excl_step) {
        const int o = t.node_owner[node];
        if (!(o == -1 || o == gid)) dd |= 0x80000000;
      }
      T.dBase[d] = t.base[node];
      T.dFree[d] = t.free_[node];
      T.dDom[d] = dd;
    }
    cnt_done = cnt;
    __syncthreads();
    for (int d = tid; d < cnt; d += nthreads) T.dAvail[d] = T.dFree[d] - T.cons[T.dSlot[d]];
    __syncthreads();

    // ---- C. sparse corrections of this step's matrix rows (fire and forget)
    {
      float* const mrow0 = b.matrix + (size_t)h.rep_off * stride - t.slab_lo;  // mrow0[node]
      for (int d = tid; d < cnt; d += nthreads) {
        const int slot = T.dSlot[d];
        const int av = T.dAvail[d];
        const bool consumed = T.cons[slot] > 0;
        float* rowp = mrow0 + T.node[slot];
        for (int p = 0; p < h.P; ++p) {
          const int count = sRole[p].count;
          if (consumed && av < sRole[p].demand) {
            for (int k = 0; k < count; ++k) rowp[(size_t)k * stride] = -INFINITY;
          } else {
            const float add = gtab_delta(T, sPair + p * RBGTOPO_MAX_GROUP_ROLES, Q, slot);
            if (add != 0.0f)
              for (int k = 0; k < count; ++k) sel_red_add_f32(rowp + (size_t)k * stride, add);
          }
          rowp += (size_t)count * stride;
        }
      }
    }

    // ---- D. exclusive domain, selection (warp p = role row p)
    int dstar = excl_step ? fixed : -1;
    if (excl_step && fixed < 0) {
      int pstar = -1;
      for (int p = 0; p < h.P; ++p)
        if (sRole[p].flags & RBGTOPO_ROLE_EXCLUSIVE) { pstar = p; break; }
      if (warp == 0) {
        int d = -1;
        if (pstar >= 0) {
          select_role_group(t, gid, excl_step, sRole[pstar], sPair + pstar * RBGTOPO_MAX_GROUP_ROLES, Q, 1, DOM_ANY, T,
                            cnt, sAcc, sAccAv, sPat, sPatAv, sList, sListAv);
          const unsigned long long top = sList[0];
          d = top ? t.domain[key_node(top)] : -1;
        }
        if (lane == 0) sDstar = d;
      }
      __syncthreads();
      dstar = sDstar;
    }
    if (warp < h.P) {
      const int p = warp;
      const bool rexcl = excl_step && (sRole[p].flags & RBGTOPO_ROLE_EXCLUSIVE);
      const int dom = rexcl ? (dstar >= 0 ? dstar : DOM_NONE) : DOM_ANY;
      int K = 0;
      for (int q = 0; q <= p; ++q) K += sRole[q].count;
      K = min(K, t.n);
      select_role_group(t, gid, excl_step, sRole[p], sPair + p * RBGTOPO_MAX_GROUP_ROLES, Q, K, dom, T, cnt,
                        sAcc + p * KS, sAccAv + p * KS, sPat + p * KS, sPatAv + p * KS, sList + p * KS,
                        sListAv + p * KS);
      b.merged[(size_t)(h.rolerow_off + p) * KS + lane] = sList[p * KS + lane];
    }
    __syncthreads();

    // ---- E. greedy from shared memory (spec §3.6); the placements stay in sTaken* for the next wave
    if (warp == 0) {
      int ntaken = 0, unplaced = 0, r = 0;
      for (int p = 0; p < h.P; ++p) {
        const int count = sRole[p].count, demand = sRole[p].demand;
        const int grole = (sRole[p].flags >> 8) & 0xFF;
        for (int c = 0; c < count; ++c, ++r) {
          int pick = -1;
          for (int k = 0; k < KS; ++k) {
            const unsigned long long key = sList[p * KS + k];
            if (key == 0) break;
            const int node = key_node(key);
            int used = 0;
            for (int i = lane; i < ntaken; i += 32)
              if (sTakenNode[i] == node) used += sTakenAmt[i];
            used = __reduce_add_sync(FULL, used);
            if (sListAv[p * KS + k] - used >= demand) {
              pick = node;
              break;
            }
          }
          if (pick >= 0) {
            if (lane == 0) {
              sTakenNode[ntaken] = pick;
              sTakenAmt[ntaken] = demand;
              sTakenRole[ntaken] = grole;
            }
            ++ntaken;
            __syncwarp();
          } else {
            ++unplaced;
          }
          if (lane == 0) b.assign[h.rep_off + r] = pick;
        }
      }
      __syncwarp();
      int status = unplaced ? RBGTOPO_PLACED_PART : RBGTOPO_PLACED_ALL;
      if (unplaced && gang) {
        status = RBGTOPO_GANG_FAILED;
        for (int i = lane; i < h.R; i += 32) b.assign[h.rep_off + i] = -1;
        ntaken = 0;
      }
      if (lane == 0) {
        b.status[step] = status;
        b.domain_out[step] = dstar;
        b.dstar[step] = dstar;
        sNew = ntaken;
        sStatus = status;
        sAny = ntaken > 0;
      }
    }
    __syncthreads();
    if (h.next_step <= 0) break;
    n_new = sNew;
    if (excl_step && dstar >= 0 && sAny) fixed = dstar;
    if (sStatus == RBGTOPO_GANG_FAILED) {  // nothing of the group is placed: its later waves report the failure
      if (warp == 0)
        for (int s2 = h.next_step; s2 > 0;) {
          const StepHdr h2 = load_hdr(b, s2);
          for (int i = lane; i < h2.R; i += 32) b.assign[h2.rep_off + i] = -1;
          if (lane == 0) {
            b.status[s2] = RBGTOPO_GANG_FAILED;
            b.domain_out[s2] = -1;
            b.dstar[s2] = -1;
          }
          s2 = h2.next_step;
        }
      break;
    }
    step = h.next_step;
    h = load_hdr(b, step);
    __syncthreads();  // sTaken*/sNew are read in A after every warp left E
  }
}

}  // namespace rbgtopo
