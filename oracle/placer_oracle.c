/*
 * placer_oracle.c — CPU ORACLE of the frozen placement spec.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rbg_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and only as the checker / the reported CPU baseline.
 *
 * PARITY UNPINNED UPSTREAM for scoring / top-K / greedy: sgl-project/rbg has no
 * such code — pkg/scheduler only manages PodGroup CRs
 * (pkg/scheduler/podgroup_manager.go:64-92) and nothing in the tree asserts
 * where pods land.  This file is therefore the literal,
 * sequential restatement of OUR spec (SURVEY.md Appendix A as frozen in
 * DESIGN.md §3), written the slow obvious way: dense role vectors, CSR rows
 * walked in storage order with an fp32 accumulator, full sort for the top-K.
 * The CUDA path computes the same thing through a different algebra (shared
 * base vector + sparse patches); bit-equality between the two is the parity
 * claim.  The pieces that ARE reference-defined and that this file consumes:
 *   - replica order inside a step = role order given (lexicographic per
 *     pkg/dependency/dependency.go:133-137), ordinal ascending
 *     (pkg/reconciler/roleinstanceset/statefulmode/stateful_instance_set_utils.go:74-76)
 *   - exclusive-topology semantics: all participating pods in ONE domain that no
 *     other group occupies (pkg/reconciler/pod_reconciler.go:172-231), per-role
 *     opt-out (pod_reconciler.go:127; api/workloads/constants/annotation.go:29)
 *   - gang = all-or-nothing (pkg/scheduler/k8s-scheduler-plugin/manager.go:131)
 * The reference-pinned arithmetic (coordination scaling etc.) lives in
 * oracle/refpinned.py and is checked against the reference's golden vectors.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 -fopenmp -shared -fPIC; the .so is built in the dev container and travels to the GPU box, so no -march=native).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define F_CAP 8
#define SELF_W 8000
#define KMAX 32
#define MAX_STEP_ROLES 8
#define MAX_GROUP_ROLES 16
#define BLOB_MAGIC 0x54474252
#define HDR_WORDS 8
#define STEP_WORDS 16

#define STEP_EXCLUSIVE 1
#define STEP_GANG 2
#define ROLE_EXCLUSIVE 1

#define ORACLE_OK 0
#define ORACLE_EINVAL -1
#define ORACLE_EINEXACT -4

typedef struct {
  int32_t n;
  int64_t e;
  const int32_t *row_ptr, *col_idx, *edge_w, *free_slots, *domain, *owner;
  int32_t n_domains;
} topo_t;

/* monotone fp32 -> u32 map, then (score desc, node asc) as one
 * descending u64 key. */
static inline uint32_t orderable_u32(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
static inline uint64_t make_key(float s, int32_t node) {
  return ((uint64_t)orderable_u32(s) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)node);
}

static int cmp_key_desc(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return (x < y) - (x > y);
}

/* One step, literally per DESIGN.md §3.  Scratch: A[n] float, S[P][n] float,
 * anchor[Q][n] int32, cons[n] int32, keys[n] u64 — allocated by the caller. */
typedef struct {
  float *A;
  float *S;        /* [MAX_STEP_ROLES][n] */
  int32_t *anchor; /* [MAX_GROUP_ROLES][n] */
  int32_t *cons;   /* [n] */
  int32_t *avail;  /* [n] */
  uint64_t *keys;  /* [n] */
} scratch_t;

static int step_place(const topo_t *t, const int32_t *blob, int64_t words,
                      const int32_t *st, scratch_t *sc, float *matrix /*nullable*/,
                      uint64_t *topk /*nullable, [P][KMAX]*/, int32_t *assign,
                      int32_t *status, int32_t *domain_out) {
  const int32_t n = t->n;
  const int32_t gid = st[0], flags = st[1], fixed_domain = st[2], P = st[3];
  const int32_t role_off = st[4], Q = st[5], pair_off = st[6];
  const int32_t n_anchors = st[7], anchor_off = st[8];
  const int32_t n_cons = st[9], cons_off = st[10], R = st[11];
  if (P < 1 || P > MAX_STEP_ROLES || Q < 0 || Q > MAX_GROUP_ROLES) return ORACLE_EINVAL;
  if (R < 1 || R > KMAX) return ORACLE_EINVAL;
  if (role_off < 0 || (int64_t)role_off + 4 * P > words) return ORACLE_EINVAL;
  if (pair_off < 0 || (int64_t)pair_off + (int64_t)P * Q > words) return ORACLE_EINVAL;
  if (n_anchors < 0 || anchor_off < 0 || (int64_t)anchor_off + 3LL * n_anchors > words) return ORACLE_EINVAL;
  if (n_cons < 0 || cons_off < 0 || (int64_t)cons_off + 2LL * n_cons > words) return ORACLE_EINVAL;
  const int32_t *roles = blob + role_off, *pair = blob + pair_off;
  const int32_t *anc = blob + anchor_off, *con = blob + cons_off;
  int32_t rsum = 0;
  for (int p = 0; p < P; ++p) {
    if (roles[4 * p] < 1 || roles[4 * p + 1] < 0 || roles[4 * p + 2] < 0) return ORACLE_EINVAL;
    rsum += roles[4 * p];
  }
  if (rsum != R) return ORACLE_EINVAL;
  const int excl_step = (flags & STEP_EXCLUSIVE) != 0;
  const int gang = (flags & STEP_GANG) != 0;

  /* A.1: dense anchor[q][m] and consumed[m] from the sparse records */
  memset(sc->anchor, 0, sizeof(int32_t) * (size_t)Q * n);
  memset(sc->cons, 0, sizeof(int32_t) * (size_t)n);
  for (int a = 0; a < n_anchors; ++a) {
    int32_t m = anc[3 * a], q = anc[3 * a + 1], c = anc[3 * a + 2];
    if (m < 0 || m >= n || q < 0 || q >= Q || c < 0) return ORACLE_EINVAL;
    sc->anchor[(size_t)q * n + m] += c;
  }
  for (int c = 0; c < n_cons; ++c) {
    int32_t m = con[2 * c], amt = con[2 * c + 1];
    if (m < 0 || m >= n || amt < 0) return ORACLE_EINVAL;
    sc->cons[m] += amt;
  }

  int inexact = 0;
  for (int p = 0; p < P; ++p) {
    const int32_t demand = roles[4 * p + 1], need = roles[4 * p + 2];
    const int role_excl = excl_step && (roles[4 * p + 3] & ROLE_EXCLUSIVE);
    /* A.2 role vector */
    for (int32_t m = 0; m < n; ++m) {
      int32_t v = 0;
      for (int q = 0; q < Q; ++q) v += pair[p * Q + q] * sc->anchor[(size_t)q * n + m];
      int32_t f = t->free_slots[m] < F_CAP ? t->free_slots[m] : F_CAP;
      v += need * f;
      sc->A[m] = (float)v;
    }
    /* A.3 score: CSR row in storage order, fp32 accumulator, then self term */
    float *S = sc->S + (size_t)p * n;
    for (int32_t i = 0; i < n; ++i) {
      float acc = 0.0f;
      for (int32_t j = t->row_ptr[i]; j < t->row_ptr[i + 1]; ++j)
        acc += (float)t->edge_w[j] * sc->A[t->col_idx[j]];
      acc += (float)SELF_W * sc->A[i];
      if (!(acc < 16777216.0f)) inexact = 1; /* A.4 exactness contract */
      int feasible = (t->free_slots[i] - sc->cons[i]) >= demand;
      if (role_excl) {
        int32_t o = t->owner[t->domain[i]];
        feasible = feasible && (o == -1 || o == gid);
      }
      S[i] = feasible ? acc : -INFINITY;
    }
  }
  if (inexact) return ORACLE_EINEXACT;

  /* dense (replica x node) matrix: replicas of a role share the role row */
  if (matrix) {
    int r = 0;
    for (int p = 0; p < P; ++p)
      for (int c = 0; c < roles[4 * p]; ++c, ++r)
        memcpy(matrix + (size_t)r * n, sc->S + (size_t)p * n, sizeof(float) * (size_t)n);
  }

  /* A.5 exclusive domain: fixed, or the domain of the best node of the first
   * participating role */
  int32_t dstar = -1;
  if (excl_step) {
    if (fixed_domain >= 0) {
      dstar = fixed_domain;
    } else {
      for (int p = 0; p < P; ++p) {
        if (!(roles[4 * p + 3] & ROLE_EXCLUSIVE)) continue;
        const float *S = sc->S + (size_t)p * n;
        uint64_t best = 0;
        int32_t bn = -1;
        for (int32_t i = 0; i < n; ++i) {
          if (S[i] == -INFINITY) continue;
          uint64_t k = make_key(S[i], i);
          if (k > best) { best = k; bn = i; }
        }
        if (bn >= 0) dstar = t->domain[bn];
        break; /* only the FIRST participating role decides */
      }
    }
  }
  *domain_out = dstar;

  /* §3.5 selection: top-K_p feasible nodes per role row, K_p = min(n, replicas of
   * the step up to and including role p): the replicas placed before role p's
   * last one can exhaust at most K_p - 1 distinct nodes, so K_p always suffices. */
  uint64_t lists[MAX_STEP_ROLES][KMAX];
  int kacc = 0;
  for (int p = 0; p < P; ++p) {
    kacc += roles[4 * p];
    const int K = kacc < n ? kacc : n;
    const int role_excl = excl_step && (roles[4 * p + 3] & ROLE_EXCLUSIVE);
    const float *S = sc->S + (size_t)p * n;
    int32_t cnt = 0;
    for (int32_t i = 0; i < n; ++i) {
      if (S[i] == -INFINITY) continue;
      if (role_excl && t->domain[i] != dstar) continue;
      sc->keys[cnt++] = make_key(S[i], i);
    }
    qsort(sc->keys, (size_t)cnt, sizeof(uint64_t), cmp_key_desc);
    for (int k = 0; k < KMAX; ++k) lists[p][k] = (k < K && k < cnt) ? sc->keys[k] : 0;
    if (topk) memcpy(topk + (size_t)p * KMAX, lists[p], sizeof(uint64_t) * KMAX);
  }

  /* A.6 greedy in replica order on a working copy of the capacity */
  for (int32_t i = 0; i < n; ++i) sc->avail[i] = t->free_slots[i] - sc->cons[i];
  int r = 0, unplaced = 0;
  for (int p = 0; p < P; ++p) {
    const int32_t demand = roles[4 * p + 1];
    for (int c = 0; c < roles[4 * p]; ++c, ++r) {
      int32_t pick = -1;
      for (int k = 0; k < KMAX; ++k) { /* K_p keys, then zeros */
        uint64_t key = lists[p][k];
        if (!key) break;
        int32_t node = (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu));
        if (sc->avail[node] >= demand) { pick = node; break; }
      }
      if (pick >= 0) sc->avail[pick] -= demand; else ++unplaced;
      assign[r] = pick;
    }
  }
  if (unplaced && gang) {
    for (int i = 0; i < R; ++i) assign[i] = -1;
    *status = 2;
  } else {
    *status = unplaced ? 1 : 0;
  }
  return ORACLE_OK;
}

static int scratch_alloc(scratch_t *sc, int32_t n) {
  sc->A = (float *)malloc(sizeof(float) * (size_t)n);
  sc->S = (float *)malloc(sizeof(float) * (size_t)n * MAX_STEP_ROLES);
  sc->anchor = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * MAX_GROUP_ROLES);
  sc->cons = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  sc->avail = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  sc->keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
  return sc->A && sc->S && sc->anchor && sc->cons && sc->avail && sc->keys;
}
static void scratch_free(scratch_t *sc) {
  free(sc->A); free(sc->S); free(sc->anchor); free(sc->cons); free(sc->avail); free(sc->keys);
}

/* Validates the snapshot the way the spec demands (symmetric CSR, sorted rows,
 * no self loops, ranges).  Returns 0 or ORACLE_EINVAL. */
int oracle_check_topology(int32_t n, int64_t e, const int32_t *row_ptr,
                          const int32_t *col_idx, const int32_t *edge_w,
                          const int32_t *free_slots, const int32_t *domain,
                          int32_t n_domains, const int32_t *owner) {
  if (n < 1 || e < 0 || row_ptr[0] != 0 || row_ptr[n] != e) return ORACLE_EINVAL;
  for (int32_t i = 0; i < n; ++i) {
    if (row_ptr[i + 1] < row_ptr[i]) return ORACLE_EINVAL;
    if (free_slots[i] < 0 || free_slots[i] > 32767) return ORACLE_EINVAL;
    if (domain[i] < 0 || domain[i] >= n_domains) return ORACLE_EINVAL;
    for (int32_t j = row_ptr[i]; j < row_ptr[i + 1]; ++j) {
      int32_t c = col_idx[j];
      if (c < 0 || c >= n || c == i) return ORACLE_EINVAL;
      if (j > row_ptr[i] && col_idx[j - 1] >= c) return ORACLE_EINVAL;
      if (edge_w[j] < 0 || edge_w[j] > 65535) return ORACLE_EINVAL;
      /* symmetric: find i in row c with the same weight */
      int32_t lo = row_ptr[c], hi = row_ptr[c + 1] - 1, found = 0;
      while (lo <= hi) {
        int32_t mid = (lo + hi) >> 1;
        if (col_idx[mid] == i) { found = (edge_w[mid] == edge_w[j]); break; }
        if (col_idx[mid] < i) lo = mid + 1; else hi = mid - 1;
      }
      if (!found) return ORACLE_EINVAL;
    }
  }
  for (int32_t d = 0; d < n_domains; ++d)
    if (owner[d] < -1) return ORACLE_EINVAL;
  return ORACLE_OK;
}

/*
 * Whole batch.  matrix: nullable [total R][n] floats; topk: nullable
 * [total role rows][KMAX] u64; assign [total R]; status/domain_out [n_steps].
 * nthreads <= 1: the canonical single-thread oracle.  nthreads > 1: steps are
 * independent (snapshot semantics) and are spread over OpenMP
 * threads — same per-step code, same results.
 */
int oracle_place(int32_t n, int64_t e, const int32_t *row_ptr, const int32_t *col_idx,
                 const int32_t *edge_w, const int32_t *free_slots, const int32_t *domain,
                 int32_t n_domains, const int32_t *owner, const int32_t *blob,
                 int64_t words, float *matrix, uint64_t *topk, int32_t *assign,
                 int32_t *status, int32_t *domain_out, int32_t nthreads) {
  if (words < HDR_WORDS || blob[0] != BLOB_MAGIC || blob[1] != 1) return ORACLE_EINVAL;
  const int32_t n_steps = blob[2];
  if (blob[3] != words || n_steps < 0) return ORACLE_EINVAL;
  if ((int64_t)HDR_WORDS + (int64_t)n_steps * STEP_WORDS > words) return ORACLE_EINVAL;
  topo_t t = {n, e, row_ptr, col_idx, edge_w, free_slots, domain, owner, n_domains};
  /* replica_off / rolerow_off must be the running prefix sums */
  int64_t racc = 0, pacc = 0;
  for (int32_t s = 0; s < n_steps; ++s) {
    const int32_t *st = blob + HDR_WORDS + (int64_t)s * STEP_WORDS;
    if (st[12] != racc || st[13] != pacc) return ORACLE_EINVAL;
    if (st[3] < 1 || st[3] > MAX_STEP_ROLES || st[11] < 1 || st[11] > KMAX) return ORACLE_EINVAL;
    racc += st[11];
    pacc += st[3];
  }
  if (blob[4] != racc || blob[5] != pacc) return ORACLE_EINVAL;

  int rc_all = ORACLE_OK;
#ifdef _OPENMP
  int nt = nthreads > 1 ? nthreads : 1;
#pragma omp parallel num_threads(nt)
#endif
  {
    scratch_t sc;
    int ok = scratch_alloc(&sc, n);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int32_t s = 0; s < n_steps; ++s) {
      const int32_t *st = blob + HDR_WORDS + (int64_t)s * STEP_WORDS;
      int rc = ok ? step_place(&t, blob, words, st, &sc,
                               matrix ? matrix + (size_t)st[12] * n : NULL,
                               topk ? topk + (size_t)st[13] * KMAX : NULL,
                               assign + st[12], status + s, domain_out + s)
                  : ORACLE_EINVAL;
      if (rc != ORACLE_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
        rc_all = rc;
      }
    }
    scratch_free(&sc);
  }
  (void)nthreads;
  return rc_all;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
