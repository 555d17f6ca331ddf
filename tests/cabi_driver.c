/*
 * cabi_driver.c — calls librbgtopo.so exactly the way the cgo shim does
 * (go/pkg/scheduler/b200topo/cgo_bridge.go): plain C, int32 arrays and sizes, one call + the
 * error fetch per helper, ten OS threads on one ctx (cgo pins each in-flight call to an OS thread;
 * --max-concurrent-reconciles defaults to 10, cmd/rbgs/main.go:140-143).  TEST INFRASTRUCTURE.
 *
 *   cabi_driver host   no device needed: the reference-pinned host arithmetic, the no-device
 *                      failure of rbgtopo_create, the thread-agnostic last_error fallback
 *   cabi_driver gpu    a synthetic 2-tier topology + a fleet of 3-role groups, placed once
 *                      sequentially and then 10 x 5 times concurrently: every result identical
 * Prints "CABI_OK <mode>" and exits 0 on success.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rbgtopo.h"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
  fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)

/* the cgo preamble helper: call + error text on the same OS thread */
static int32_t go_place_groups(rbgtopo_ctx* ctx, const int32_t* groups, int64_t words, int32_t* assign, int32_t* status,
                               int32_t* domain, char* err, int errlen) {
  int32_t rc = rbgtopo_place_groups(ctx, groups, words, assign, status, domain);
  if (rc != RBGTOPO_OK) rbgtopo_last_error(ctx, err, errlen); else err[0] = 0;
  return rc;
}

/* ------------------------------------------------------------------ host */
static void* other_thread_reads_error(void* arg) {
  char* buf = (char*)arg;
  rbgtopo_last_error(NULL, buf, 256); /* this thread never failed: the library-wide fallback answers */
  return NULL;
}

static int run_host(void) {
  CHECK(rbgtopo_abi_version() == RBGTOPO_ABI_VERSION, "abi version");
  /* GetGroupSize, api/workloads/v1alpha2/helper.go:50-65: mooncake pd = 7 pods, an LWS role counts size x replicas */
  int32_t rep[5] = {1, 3, 1, 1, 1}, lws[5] = {0, 0, 0, 0, 0};
  CHECK(rbgtopo_group_size(5, rep, lws) == 7, "group size");
  int32_t rep2[2] = {2, 3}, lws2[2] = {4, 0};
  CHECK(rbgtopo_group_size(2, rep2, lws2) == 11, "group size with LWS");
  /* dependencyOrder, pkg/dependency/dependency_test.go:37-121: a -> b -> c  =>  [[c],[b],[a]] */
  const char* names[3] = {"a", "b", "c"};
  int32_t dep_off[4] = {0, 1, 2, 2}, dep_idx[2] = {1, 2}, level[3], order[3];
  CHECK(rbgtopo_dependency_levels(3, names, dep_off, dep_idx, level, order) == 3, "levels");
  CHECK(level[0] == 2 && level[1] == 1 && level[2] == 0 && order[0] == 2 && order[2] == 0, "level order");
  int32_t cyc_off[3] = {0, 1, 2}, cyc_idx[2] = {1, 0};
  CHECK(rbgtopo_dependency_levels(2, names, cyc_off, cyc_idx, level, order) == RBGTOPO_EINVAL, "cycle");
  /* parsePercentage, scaler_test.go:520-597 */
  double v = 0;
  CHECK(rbgtopo_parse_percentage(" 5% ", &v) == 0 && v == 0.05, "percentage");
  CHECK(rbgtopo_parse_percentage("150%", &v) != 0, "percentage range");
  /* CalculateTargetReplicas, scaler_test.go:100-518 first case: 5 %% of (300, 100) from zero => 15 / 5 */
  int32_t desired[2] = {300, 100}, zero[2] = {0, 0}, target[2];
  CHECK(rbgtopo_calculate_target_replicas(0.05, 0, 2, desired, zero, zero, zero, target) == 0, "target rc");
  CHECK(target[0] == 15 && target[1] == 5, "targets %d %d", target[0], target[1]);
  /* calculateCoordinationUpdatedReplicasBound, rolebasedgroup_controller_test.go:1283-1377: 1 %%, 20/200, 100 -> [9, 11] */
  int32_t lo = 0, hi = 0;
  CHECK(rbgtopo_updated_replicas_bound(1, 20, 200, 100, &lo, &hi) == 0 && lo == 9 && hi == 11, "bound %d %d", lo, hi);
  /* GetScaledValueFromIntOrPercent, intstr.go:181-197 */
  CHECK(rbgtopo_scaled_value(1, 25, 10, 1) == 3 && rbgtopo_scaled_value(1, 25, 10, 0) == 2, "scaled value");
  /* CalculatePartitionReplicas / ParseIntStrAsNonZero, pkg/utils/utils.go:139-162,177-185 */
  int32_t out = -1;
  CHECK(rbgtopo_partition_replicas(1, 1, 30, 10, &out) == 0 && out == 3, "partition %d", out);
  CHECK(rbgtopo_partition_replicas(0, 0, 0, 10, &out) == 0 && out == 0, "nil partition");
  CHECK(rbgtopo_intstr_non_zero(1, 5, 10, &out) == 0 && out == 1, "non zero %d", out);

  /* no device in this mode's environment is the common case; either way the call must not abort */
  rbgtopo_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.world = 1;
  rbgtopo_ctx* ctx = NULL;
  int32_t rc = rbgtopo_create(&cfg, &ctx);
  if (rc != RBGTOPO_OK) {
    char mine[256], theirs[256];
    CHECK(rc == RBGTOPO_ENODEVICE || rc == RBGTOPO_ECUDA, "create rc %d", rc);
    CHECK(ctx == NULL, "ctx must stay NULL on failure");
    CHECK(rbgtopo_last_error(NULL, mine, sizeof mine) > 0 && strlen(mine) > 0, "error text");
    pthread_t th; /* a goroutine that migrated to another OS thread still gets the text */
    pthread_create(&th, NULL, other_thread_reads_error, theirs);
    pthread_join(th, NULL);
    CHECK(strcmp(mine, theirs) == 0, "fallback text: '%s' vs '%s'", mine, theirs);
  } else {
    rbgtopo_destroy(ctx);
  }
  /* bad arguments come back as codes, never as crashes */
  CHECK(rbgtopo_place_groups(NULL, NULL, 0, NULL, NULL, NULL) == RBGTOPO_EINVAL, "null ctx");
  printf("CABI_OK host\n");
  return 0;
}

/* ------------------------------------------------------------------- gpu */
#define NN 4096
#define NG 48
typedef struct {
  rbgtopo_ctx* ctx;
  const int32_t* blob;
  int64_t words;
  int n_pending;
  const int32_t* want_assign;
  const int32_t* want_status;
  int bad;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  int32_t* assign = (int32_t*)malloc(sizeof(int32_t) * (size_t)j->n_pending);
  int32_t status[NG], domain[NG];
  char err[256];
  for (int it = 0; it < 5; ++it) {
    int32_t rc = go_place_groups(j->ctx, j->blob, j->words, assign, status, domain, err, sizeof err);
    if (rc != RBGTOPO_OK || memcmp(assign, j->want_assign, sizeof(int32_t) * (size_t)j->n_pending) != 0 ||
        memcmp(status, j->want_status, sizeof status) != 0)
      j->bad++;
  }
  free(assign);
  return NULL;
}

static int run_gpu(void) {
  rbgtopo_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.world = 1;
  rbgtopo_ctx* ctx = NULL;
  char err[256];
  int32_t rc = rbgtopo_create(&cfg, &ctx);
  if (rc != RBGTOPO_OK) {
    rbgtopo_last_error(NULL, err, sizeof err);
    CHECK(0, "rbgtopo_create: %d %s", rc, err);
  }
  /* topology: NVLink cliques of 8 (weight 1000) + a ring across domains (weight 10), symmetric, sorted rows */
  static int32_t row_ptr[NN + 1], col[NN * 9], w[NN * 9], free_slots[NN], domain[NN], owner[NN / 8];
  int64_t e = 0;
  for (int i = 0; i < NN; ++i) {
    row_ptr[i] = (int32_t)e;
    int nb[9], nw[9], k = 0;
    for (int o = 0; o < 8; ++o) {
      int p = (i / 8) * 8 + o;
      if (p != i) { nb[k] = p; nw[k++] = 1000; }
    }
    nb[k] = (i + 8) % NN; nw[k++] = 10;
    nb[k] = (i + NN - 8) % NN; nw[k++] = 10;
    for (int a = 0; a < k; ++a)   /* insertion sort by column */
      for (int b = a + 1; b < k; ++b)
        if (nb[b] < nb[a]) { int t = nb[a]; nb[a] = nb[b]; nb[b] = t; t = nw[a]; nw[a] = nw[b]; nw[b] = t; }
    for (int a = 0; a < k; ++a) { col[e] = nb[a]; w[e] = nw[a]; ++e; }
    free_slots[i] = (int32_t)((i * 2654435761u >> 7) % 9);
    domain[i] = i / 8;
  }
  row_ptr[NN] = (int32_t)e;
  for (int d = 0; d < NN / 8; ++d) owner[d] = -1;
  rc = rbgtopo_set_topology(ctx, NN, e, row_ptr, col, w, free_slots, domain, NN / 8, owner, 1);
  if (rc != RBGTOPO_OK) { rbgtopo_last_error(ctx, err, sizeof err); CHECK(0, "set_topology: %s", err); }

  /* GROUPS blob: NG groups, roles (level, pending, demand, flags): a(0,1,1) | b(1,3,1), c(1,2,1); pair = all ones */
  const int q = 3, per = 4 * q + q * q + 3;
  const int words = RBGTOPO_HDR_WORDS + NG * RBGTOPO_GROUP_WORDS + NG * per;
  int32_t* blob = (int32_t*)calloc((size_t)words, sizeof(int32_t));
  blob[0] = RBGTOPO_GROUPS_MAGIC; blob[1] = RBGTOPO_ABI_VERSION; blob[2] = NG; blob[3] = words; blob[4] = NG * 6;
  int off = RBGTOPO_HDR_WORDS + NG * RBGTOPO_GROUP_WORDS;
  for (int g = 0; g < NG; ++g) {
    int32_t* rec = blob + RBGTOPO_HDR_WORDS + g * RBGTOPO_GROUP_WORDS;
    rec[0] = g; rec[1] = (g % 4 == 0) ? RBGTOPO_STEP_GANG : 0; rec[2] = -1; rec[3] = q;
    rec[4] = off;
    const int32_t roles[12] = {0, 1, 1, RBGTOPO_ROLE_EXCLUSIVE, 1, 3, 1, RBGTOPO_ROLE_EXCLUSIVE, 1, 2, 1, RBGTOPO_ROLE_EXCLUSIVE};
    memcpy(blob + off, roles, sizeof roles); off += 12;
    rec[5] = off;
    for (int i = 0; i < q * q; ++i) blob[off++] = 1;
    rec[6] = 1; rec[7] = off;
    blob[off++] = (g * 83) % NN; blob[off++] = 0; blob[off++] = 1; /* one scheduled pod of role a */
    rec[8] = g * 6; rec[9] = 6;
  }
  CHECK(off == words, "blob size");
  const int n_pending = NG * 6;
  int32_t* want = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_pending);
  int32_t want_status[NG], want_domain[NG];
  rc = go_place_groups(ctx, blob, words, want, want_status, want_domain, err, sizeof err);
  CHECK(rc == RBGTOPO_OK, "place_groups: %d %s", rc, err);
  int placed = 0;
  for (int i = 0; i < n_pending; ++i) {
    CHECK(want[i] >= -1 && want[i] < NN, "assign[%d] = %d", i, want[i]);
    placed += want[i] >= 0;
    if (want[i] >= 0) CHECK(free_slots[want[i]] >= 1, "replica %d on a full node", i);
  }
  CHECK(placed > n_pending / 2, "only %d of %d placed", placed, n_pending);
  /* malformed input: a code and a message, the ctx stays usable */
  blob[RBGTOPO_HDR_WORDS + 1] = 64; /* unknown flag bit */
  rc = go_place_groups(ctx, blob, words, want_domain, want_domain, want_domain, err, sizeof err);
  CHECK(rc == RBGTOPO_EINVAL && strstr(err, "flags"), "unknown flags: %d '%s'", rc, err);
  blob[RBGTOPO_HDR_WORDS + 1] = RBGTOPO_STEP_GANG;

  pthread_t th[10];
  job_t jobs[10];
  for (int t = 0; t < 10; ++t) {
    jobs[t] = (job_t){ctx, blob, words, n_pending, want, want_status, 0};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int bad = 0;
  for (int t = 0; t < 10; ++t) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
  CHECK(bad == 0, "%d concurrent calls differ from the sequential result", bad);
  uint64_t gen = 0; int64_t calls = 0, scores = 0, launches = 0;
  CHECK(rbgtopo_stats(ctx, &gen, &calls, &scores, &launches) == 0 && gen == 1 && calls >= 51 && launches > 0, "stats");
  rbgtopo_destroy(ctx);
  free(blob); free(want);
  printf("CABI_OK gpu\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "gpu") == 0) return run_gpu();
  return run_host();
}
