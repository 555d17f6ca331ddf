/*
 * placer_fast.c — CPU variant of the oracle with the SAME ALGEBRA as the CUDA path.
 *
 * TEST INFRASTRUCTURE ONLY (same status as placer_oracle.c: only tests/ and bench.py's
 * cpu_baseline / --impl reference legs may call it).  PARITY UNPINNED UPSTREAM: sgl-project/rbg has
 * no scoring / top-K / greedy code; this is OUR spec.
 *
 * Why it exists (VERDICT r1 "measurement hygiene"): placer_oracle.c is the literal, slow-obvious
 * restatement (dense role vectors, a full SpMV per role row, qsort of every feasible key), while the
 * GPU computes  S = need * base + sparse anchor terms   and selects from the
 * per-snapshot background order plus the few patched nodes (§4.3).  Timing only the literal oracle
 * makes the GPU / CPU ratio flatter the algebra, not the hardware.  This file is the honest CPU
 * baseline: one SpMV per SNAPSHOT (base = W * min(free, 8)), one sort per snapshot (background
 * order), and per step a dense row = one multiply per score, sparse patches from a small hash
 * table, top-K from (patched nodes) U (first K feasible unpatched nodes of the order).  Every term
 * is an exact integer below 2^24 (spec §3.4), so the results equal placer_oracle.c bit for bit —
 * tests/test_oracle_fast.py checks matrix bits, top-K keys, assignment, status and domain.
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

static inline uint32_t orderable_u32(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
static inline uint64_t make_key(float s, int32_t node) {
  return ((uint64_t)orderable_u32(s) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)node);
}
static inline int32_t key_node(uint64_t k) { return (int32_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu)); }
static int cmp_key_desc(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return (x < y) - (x > y);
}

/* ---- per-snapshot state, cached across calls (bench.py calls once per wave on one snapshot) */
typedef struct {
  const int32_t *row_ptr, *col_idx, *edge_w;
  int32_t n;
  int64_t e;
  uint64_t free_sum;   /* checksum of free[] the cache was built for */
  float *base;         /* W * min(free, F) */
  float max_base;
  int32_t *order;      /* nodes by key(base, node) descending */
} snap_t;
static snap_t g_snap;

static uint64_t checksum(const int32_t *v, int32_t n) {
  uint64_t h = 1469598103934665603ull;
  for (int32_t i = 0; i < n; ++i) h = (h ^ (uint32_t)v[i]) * 1099511628211ull;
  return h;
}

static int snapshot_prepare(int32_t n, int64_t e, const int32_t *row_ptr, const int32_t *col_idx, const int32_t *edge_w,
                            const int32_t *free_slots, int nthreads) {
  const uint64_t fs = checksum(free_slots, n);
  if (g_snap.base && g_snap.row_ptr == row_ptr && g_snap.col_idx == col_idx && g_snap.edge_w == edge_w && g_snap.n == n &&
      g_snap.e == e && g_snap.free_sum == fs)
    return ORACLE_OK;
  free(g_snap.base);
  free(g_snap.order);
  g_snap.base = (float *)malloc(sizeof(float) * (size_t)n);
  g_snap.order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
  if (!g_snap.base || !g_snap.order || !keys) return ORACLE_EINVAL;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (int32_t i = 0; i < n; ++i) {
    float acc = 0.0f;
    for (int32_t j = row_ptr[i]; j < row_ptr[i + 1]; ++j) {
      int32_t f = free_slots[col_idx[j]];
      acc += (float)edge_w[j] * (float)(f < F_CAP ? f : F_CAP);
    }
    int32_t f = free_slots[i];
    acc += (float)SELF_W * (float)(f < F_CAP ? f : F_CAP);
    g_snap.base[i] = acc;
    keys[i] = make_key(acc, i);
  }
  qsort(keys, (size_t)n, sizeof(uint64_t), cmp_key_desc);
  for (int32_t i = 0; i < n; ++i) g_snap.order[i] = key_node(keys[i]);
  g_snap.max_base = n > 0 ? g_snap.base[g_snap.order[0]] : 0.0f;
  free(keys);
  g_snap.row_ptr = row_ptr; g_snap.col_idx = col_idx; g_snap.edge_w = edge_w;
  g_snap.n = n; g_snap.e = e; g_snap.free_sum = fs;
  (void)nthreads;
  return ORACLE_OK;
}

/* ---- per-thread patch table: open addressing on the node id */
typedef struct {
  int32_t cap, mask, cnt;
  int32_t *node;            /* [cap], -1 = empty */
  int32_t *slots;           /* [cnt] occupied positions, insertion order */
  int32_t *cons;            /* [cap] */
  float *delta;             /* [cap][MAX_STEP_ROLES] */
} tab_t;

static int tab_reserve(tab_t *T, int32_t want) {
  int32_t cap = 64;
  while (cap < 2 * want) cap <<= 1;
  if (cap > T->cap) {
    free(T->node); free(T->slots); free(T->cons); free(T->delta);
    T->node = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    T->slots = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    T->cons = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    T->delta = (float *)malloc(sizeof(float) * (size_t)cap * MAX_STEP_ROLES);
    if (!T->node || !T->slots || !T->cons || !T->delta) return 0;
    T->cap = cap;
    for (int32_t i = 0; i < cap; ++i) T->node[i] = -1;
  }
  T->mask = T->cap - 1;
  T->cnt = 0;
  return 1;
}
static inline int32_t tab_slot(tab_t *T, int32_t n, int insert) {
  int32_t h = (int32_t)(((uint32_t)n * 2654435761u) >> 12) & T->mask;
  while (1) {
    if (T->node[h] == n) return h;
    if (T->node[h] == -1) {
      if (!insert) return -1;
      T->node[h] = n;
      T->cons[h] = 0;
      memset(T->delta + (size_t)h * MAX_STEP_ROLES, 0, sizeof(float) * MAX_STEP_ROLES);
      T->slots[T->cnt++] = h;
      return h;
    }
    h = (h + 1) & T->mask;
  }
}
static void tab_clear(tab_t *T) {
  for (int32_t i = 0; i < T->cnt; ++i) T->node[T->slots[i]] = -1;
  T->cnt = 0;
}

typedef struct { tab_t T; uint64_t *keys; int32_t keys_cap; } scratch_t;

static int step_fast(int32_t n, const int32_t *row_ptr, const int32_t *col_idx, const int32_t *edge_w,
                     const int32_t *free_slots, const int32_t *domain, const int32_t *owner, const int32_t *blob,
                     int64_t words, const int32_t *st, scratch_t *sc, float *matrix, uint64_t *topk, int32_t *assign,
                     int32_t *status, int32_t *domain_out) {
  const int32_t gid = st[0], flags = st[1], fixed_domain = st[2], P = st[3];
  const int32_t role_off = st[4], Q = st[5], pair_off = st[6];
  const int32_t n_anchors = st[7], anchor_off = st[8], n_cons = st[9], cons_off = st[10], R = st[11];
  if (P < 1 || P > MAX_STEP_ROLES || Q < 0 || Q > MAX_GROUP_ROLES || R < 1 || R > KMAX) return ORACLE_EINVAL;
  if (role_off < 0 || (int64_t)role_off + 4 * P > words || pair_off < 0 || (int64_t)pair_off + (int64_t)P * Q > words ||
      n_anchors < 0 || anchor_off < 0 || (int64_t)anchor_off + 3LL * n_anchors > words || n_cons < 0 || cons_off < 0 ||
      (int64_t)cons_off + 2LL * n_cons > words)
    return ORACLE_EINVAL;
  const int32_t *roles = blob + role_off, *pair = blob + pair_off, *anc = blob + anchor_off, *con = blob + cons_off;
  const int excl_step = (flags & STEP_EXCLUSIVE) != 0, gang = (flags & STEP_GANG) != 0;
  const float *base = g_snap.base;
  tab_t *T = &sc->T;

  /* patched nodes: closed neighbourhoods of the anchor pods + nodes with consumed capacity */
  int64_t want = n_cons;
  for (int a = 0; a < n_anchors; ++a) {
    const int32_t m = anc[3 * a];
    if (m < 0 || m >= n || anc[3 * a + 1] < 0 || anc[3 * a + 1] >= Q || anc[3 * a + 2] < 0) return ORACLE_EINVAL;
    want += row_ptr[m + 1] - row_ptr[m] + 1;
  }
  if (!tab_reserve(T, (int32_t)(want < 16 ? 16 : want))) return ORACLE_EINVAL;
  for (int a = 0; a < n_anchors; ++a) {
    const int32_t m = anc[3 * a], q = anc[3 * a + 1], c = anc[3 * a + 2];
    if (c == 0) continue;
    for (int32_t j = row_ptr[m]; j <= row_ptr[m + 1]; ++j) {   /* j == end stands for the anchor's own node */
      const int32_t nn = j < row_ptr[m + 1] ? col_idx[j] : m;
      const int32_t wv = (j < row_ptr[m + 1] ? edge_w[j] : SELF_W) * c;
      float *d = T->delta + (size_t)tab_slot(T, nn, 1) * MAX_STEP_ROLES;
      for (int p = 0; p < P; ++p) d[p] += (float)(pair[p * Q + q] * wv);
    }
  }
  for (int c = 0; c < n_cons; ++c) {
    const int32_t m = con[2 * c], amt = con[2 * c + 1];
    if (m < 0 || m >= n || amt < 0) { tab_clear(T); return ORACLE_EINVAL; }
    T->cons[tab_slot(T, m, 1)] += amt;
  }

  /* dense rows: one multiply per score, then the sparse patches */
  int inexact = 0;
  for (int p = 0; p < P; ++p)   /* spec §3.4 for the background scores (the patched ones are checked where computed) */
    if (!((float)roles[4 * p + 2] * g_snap.max_base < 16777216.0f)) inexact = 1;
  if (matrix) {
    int r = 0;
    for (int p = 0; p < P; ++p) {
      const int32_t count = roles[4 * p], demand = roles[4 * p + 1];
      const float need = (float)roles[4 * p + 2];
      const int role_excl = excl_step && (roles[4 * p + 3] & ROLE_EXCLUSIVE);
      float *row = matrix + (size_t)r * n;
      if (!role_excl) {
        for (int32_t i = 0; i < n; ++i) row[i] = free_slots[i] >= demand ? need * base[i] : -INFINITY;
      } else {
        for (int32_t i = 0; i < n; ++i) {
          const int32_t o = owner[domain[i]];
          row[i] = (free_slots[i] >= demand && (o == -1 || o == gid)) ? need * base[i] : -INFINITY;
        }
      }
      for (int32_t k = 0; k < T->cnt; ++k) {
        const int32_t h = T->slots[k], i = T->node[h];
        int feasible = free_slots[i] - T->cons[h] >= demand;
        if (role_excl) {
          const int32_t o = owner[domain[i]];
          feasible = feasible && (o == -1 || o == gid);
        }
        const float s = need * base[i] + T->delta[(size_t)h * MAX_STEP_ROLES + p];
        if (!(s < 16777216.0f)) inexact = 1;
        row[i] = feasible ? s : -INFINITY;
      }
      for (int c = 1; c < count; ++c) memcpy(row + (size_t)c * n, row, sizeof(float) * (size_t)n);
      r += count;
    }
  }

  /* selection helper: top-K of role row p restricted to `dom` (-2 = any, -1 = nothing) */
  uint64_t lists[MAX_STEP_ROLES][KMAX];
#define SCORE_OF(h, i, p) (need * base[i] + T->delta[(size_t)(h) * MAX_STEP_ROLES + (p)])
  int32_t dstar = -1;
  for (int pass = 0; pass < 2; ++pass) {   /* pass 0: D* of the first participating role; pass 1: every role */
    if (pass == 0 && !(excl_step && fixed_domain < 0)) {
      if (excl_step) dstar = fixed_domain;
      continue;
    }
    int kacc = 0;
    for (int p = 0; p < P; ++p) {
      const int32_t demand = roles[4 * p + 1], need_i = roles[4 * p + 2];
      const float need = (float)need_i;
      const int role_excl = excl_step && (roles[4 * p + 3] & ROLE_EXCLUSIVE);
      kacc += roles[4 * p];
      int K = kacc < n ? kacc : n;
      int32_t dom = -2;
      if (pass == 0) {
        if (!(roles[4 * p + 3] & ROLE_EXCLUSIVE)) continue;
        K = 1;
      } else if (role_excl) {
        dom = dstar;   /* -1: no feasible domain -> empty list */
      }
      int32_t cnt = 0;
      if (dom != -1) {
        if (sc->keys_cap < T->cnt + KMAX) {
          free(sc->keys);
          sc->keys_cap = 2 * (T->cnt + KMAX);
          sc->keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)sc->keys_cap);
          if (!sc->keys) { tab_clear(T); return ORACLE_EINVAL; }
        }
        for (int32_t k = 0; k < T->cnt; ++k) {     /* (a) patched nodes */
          const int32_t h = T->slots[k], i = T->node[h];
          if (free_slots[i] - T->cons[h] < demand) continue;
          if (role_excl) {
            const int32_t o = owner[domain[i]];
            if (!(o == -1 || o == gid)) continue;
          }
          if (dom >= 0 && domain[i] != dom) continue;
          const float s = SCORE_OF(h, i, p);
          if (!(s < 16777216.0f)) inexact = 1;
          sc->keys[cnt++] = make_key(s, i);
        }
        int got = 0;                                 /* (b) first K feasible unpatched nodes of the order */
        for (int32_t pos = 0; pos < n && got < K; ++pos) {
          const int32_t i = need_i > 0 ? g_snap.order[pos] : pos;   /* need == 0: every background score is 0 */
          if (free_slots[i] < demand) continue;
          if (role_excl) {
            const int32_t o = owner[domain[i]];
            if (!(o == -1 || o == gid)) continue;
          }
          if (dom >= 0 && domain[i] != dom) continue;
          if (tab_slot(T, i, 0) >= 0) continue;
          sc->keys[cnt++] = make_key(need * base[i], i);
          ++got;
        }
        qsort(sc->keys, (size_t)cnt, sizeof(uint64_t), cmp_key_desc);
      }
      if (pass == 0) {
        dstar = cnt > 0 ? domain[key_node(sc->keys[0])] : -1;
        break;   /* only the FIRST participating role decides */
      }
      for (int k = 0; k < KMAX; ++k) lists[p][k] = (k < K && k < cnt) ? sc->keys[k] : 0;
      if (topk) memcpy(topk + (size_t)p * KMAX, lists[p], sizeof(uint64_t) * KMAX);
    }
  }
#undef SCORE_OF
  *domain_out = dstar;
  if (inexact) { tab_clear(T); return ORACLE_EINEXACT; }

  /* greedy in replica order; capacity taken in this step lives in a tiny list */
  int32_t tnode[KMAX], tamt[KMAX];
  int ntaken = 0, unplaced = 0, r = 0;
  for (int p = 0; p < P; ++p) {
    const int32_t demand = roles[4 * p + 1];
    for (int c = 0; c < roles[4 * p]; ++c, ++r) {
      int32_t pick = -1;
      for (int k = 0; k < KMAX; ++k) {
        const uint64_t key = lists[p][k];
        if (!key) break;
        const int32_t node = key_node(key);
        const int32_t h = tab_slot(T, node, 0);
        int32_t avail = free_slots[node] - (h >= 0 ? T->cons[h] : 0);
        for (int i = 0; i < ntaken; ++i)
          if (tnode[i] == node) avail -= tamt[i];
        if (avail >= demand) { pick = node; break; }
      }
      if (pick >= 0) { tnode[ntaken] = pick; tamt[ntaken++] = demand; } else ++unplaced;
      assign[r] = pick;
    }
  }
  if (unplaced && gang) {
    for (int i = 0; i < R; ++i) assign[i] = -1;
    *status = 2;
  } else {
    *status = unplaced ? 1 : 0;
  }
  tab_clear(T);
  return ORACLE_OK;
}

int oracle_place_fast(int32_t n, int64_t e, const int32_t *row_ptr, const int32_t *col_idx, const int32_t *edge_w,
                      const int32_t *free_slots, const int32_t *domain, int32_t n_domains, const int32_t *owner,
                      const int32_t *blob, int64_t words, float *matrix, uint64_t *topk, int32_t *assign,
                      int32_t *status, int32_t *domain_out, int32_t nthreads) {
  if (words < HDR_WORDS || blob[0] != BLOB_MAGIC || blob[1] != 1) return ORACLE_EINVAL;
  const int32_t n_steps = blob[2];
  if (blob[3] != words || n_steps < 0 || (int64_t)HDR_WORDS + (int64_t)n_steps * STEP_WORDS > words) return ORACLE_EINVAL;
  int64_t racc = 0, pacc = 0;
  for (int32_t s = 0; s < n_steps; ++s) {
    const int32_t *st = blob + HDR_WORDS + (int64_t)s * STEP_WORDS;
    if (st[12] != racc || st[13] != pacc || st[3] < 1 || st[3] > MAX_STEP_ROLES || st[11] < 1 || st[11] > KMAX)
      return ORACLE_EINVAL;
    racc += st[11];
    pacc += st[3];
  }
  if (blob[4] != racc || blob[5] != pacc) return ORACLE_EINVAL;
  int rc_all = snapshot_prepare(n, e, row_ptr, col_idx, edge_w, free_slots, nthreads);
  if (rc_all != ORACLE_OK) return rc_all;
  (void)n_domains;
#ifdef _OPENMP
  int nt = nthreads > 1 ? nthreads : 1;
#pragma omp parallel num_threads(nt)
#endif
  {
    static __thread scratch_t sc;   /* keeps its buffers across calls */
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int32_t s = 0; s < n_steps; ++s) {
      const int32_t *st = blob + HDR_WORDS + (int64_t)s * STEP_WORDS;
      int rc = step_fast(n, row_ptr, col_idx, edge_w, free_slots, domain, owner, blob, words, st, &sc,
                         matrix ? matrix + (size_t)st[12] * n : NULL, topk ? topk + (size_t)st[13] * KMAX : NULL,
                         assign + st[12], status + s, domain_out + s);
      if (rc != ORACLE_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
        rc_all = rc;
      }
    }
  }
  return rc_all;
}
