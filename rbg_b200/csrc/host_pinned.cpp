// host_pinned.cpp — product-side C mirrors of the reference-defined arithmetic the
// host needs to build placement steps.  In production the
// Go shim calls the controller's own Go functions; a C/C++/Python host above the
// ABI calls these.  Each function follows the cited Go code; all are pinned to
// the reference's golden tables by tests/test_host_pinned.py (through the C ABI)
// and cross-checked against oracle/refpinned.py on sweeps.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/rbgtopo.h"

extern "C" {

// RoleBasedGroup.GetGroupSize — api/workloads/v1alpha2/helper.go:50-65
int32_t rbgtopo_group_size(int32_t n_roles, const int32_t* replicas, const int32_t* lws_size) {
  if (n_roles < 0 || (n_roles && !replicas)) return RBGTOPO_EINVAL;
  int64_t ret = 0;
  for (int i = 0; i < n_roles; ++i) {
    const int32_t sz = (lws_size && lws_size[i] > 0) ? lws_size[i] : 1;
    ret += (int64_t)sz * replicas[i];
  }
  return (int32_t)ret;
}

// dependencyOrder — pkg/dependency/dependency.go:129-205: names sorted, DFS level
// = 1 + max(level of dependencies), cycle -> error.
int32_t rbgtopo_dependency_levels(int32_t n, const char* const* names, const int32_t* dep_off,
                                  const int32_t* dep_idx, int32_t* level_of, int32_t* order) {
  if (n < 0 || (n && (!names || !dep_off || !level_of || !order))) return RBGTOPO_EINVAL;
  std::vector<int> by_name(n);
  std::iota(by_name.begin(), by_name.end(), 0);
  std::sort(by_name.begin(), by_name.end(),
            [&](int a, int b) { return strcmp(names[a], names[b]) < 0; });
  std::vector<int> st(n, -2);  // -2 not started, -1 in progress, >= 0 level
  // iterative DFS with an explicit stack (role, next dependency position)
  for (int root : by_name) {
    if (st[root] != -2) continue;
    std::vector<std::pair<int, int>> stack;
    stack.push_back({root, dep_off[root]});
    st[root] = -1;
    std::vector<int> mx(n, 0);
    while (!stack.empty()) {
      auto& top = stack.back();
      const int r = top.first;
      if (top.second < dep_off[r + 1]) {
        const int d = dep_idx[top.second++];
        if (d < 0 || d >= n) return RBGTOPO_EINVAL;
        if (st[d] >= 0) {
          mx[r] = std::max(mx[r], st[d] + 1);
        } else if (st[d] == -1) {
          return RBGTOPO_EINVAL;  // cycle
        } else {
          st[d] = -1;
          stack.push_back({d, dep_off[d]});
        }
      } else {
        st[r] = mx[r];
        stack.pop_back();
        if (!stack.empty()) {
          const int parent = stack.back().first;
          mx[parent] = std::max(mx[parent], st[r] + 1);
        }
      }
    }
  }
  int levels = 0;
  for (int i = 0; i < n; ++i) {
    level_of[i] = st[i];
    levels = std::max(levels, st[i] + 1);
  }
  std::stable_sort(by_name.begin(), by_name.end(), [&](int a, int b) { return st[a] < st[b]; });
  for (int i = 0; i < n; ++i) order[i] = by_name[i];
  return n ? levels : 0;
}

// parsePercentage — pkg/coordination/coordinationscaling/scaler.go:253-270
int32_t rbgtopo_parse_percentage(const char* s, double* out) {
  if (!s || !out) return RBGTOPO_EINVAL;
  std::string t(s);
  const char* ws = " \t\n\v\f\r";
  size_t b = t.find_first_not_of(ws);
  if (b == std::string::npos) return RBGTOPO_EINVAL;
  size_t e = t.find_last_not_of(ws);
  t = t.substr(b, e - b + 1);
  if (t.empty() || t.back() != '%') return RBGTOPO_EINVAL;
  t.pop_back();
  if (t.empty() || isspace((unsigned char)t.front()) || t.find('_') != std::string::npos) return RBGTOPO_EINVAL;
  // strconv.ParseFloat accepts decimal / exponent forms; reject what strtod adds
  for (char ch : t)
    if (!(isdigit((unsigned char)ch) || ch == '.' || ch == 'e' || ch == 'E' || ch == '+' || ch == '-'))
      return RBGTOPO_EINVAL;
  errno = 0;
  char* end = nullptr;
  double v = strtod(t.c_str(), &end);
  if (end == t.c_str() || *end != 0) return RBGTOPO_EINVAL;
  if (v < 0 || v > 100) return RBGTOPO_EINVAL;
  *out = v / 100.0;
  return RBGTOPO_OK;
}

// CalculateTargetReplicas — scaler.go:70-172 with canProceedToNextBatch :192-242.
// progression: 0 = unset (the Go switch matches nothing: no gate), 1 =
// OrderScheduled, 2 = OrderReady.
int32_t rbgtopo_calculate_target_replicas(double max_skew, int32_t progression, int32_t n,
                                          const int32_t* desired, const int32_t* current,
                                          const int32_t* scheduled, const int32_t* ready,
                                          int32_t* target) {
  if (n <= 0 || !desired || !current || !scheduled || !ready || !target) return RBGTOPO_EINVAL;
  bool all_done = true;
  for (int i = 0; i < n; ++i)
    if (current[i] < desired[i]) { all_done = false; break; }
  bool proceed = true;
  if (!all_done) {
    for (int i = 0; i < n; ++i) {
      if (current[i] >= desired[i] || current[i] == 0) continue;
      if (progression == 1 && scheduled[i] < current[i]) { proceed = false; break; }
      if (progression == 2 && ready[i] < current[i]) { proceed = false; break; }
    }
  }
  if (!proceed) {
    for (int i = 0; i < n; ++i) target[i] = current[i];
    return RBGTOPO_OK;
  }
  std::vector<double> prog(n);
  for (int i = 0; i < n; ++i) {
    if (desired[i] == 0) prog[i] = current[i] == 0 ? 1.0 : 0.0;
    else prog[i] = (double)current[i] / (double)desired[i];
  }
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return prog[a] < prog[b]; });
  double min_progress = prog[idx[0]];
  for (int k : idx)
    if (current[k] < desired[k]) { min_progress = prog[k]; break; }
  const double max_allowed = min_progress + max_skew;
  for (int i = 0; i < n; ++i) {
    if (current[i] >= desired[i]) { target[i] = desired[i]; continue; }
    if (prog[i] >= max_allowed) { target[i] = current[i]; continue; }
    int32_t t = (int32_t)std::ceil(max_allowed * (double)desired[i]);
    if (t > desired[i]) t = desired[i];
    if (t <= current[i] && current[i] < desired[i]) t = current[i] + 1;
    target[i] = t;
  }
  return RBGTOPO_OK;
}

// GetScaledValueFromIntOrPercent — vendor/k8s.io/apimachinery/pkg/util/intstr/intstr.go:181-197
int32_t rbgtopo_scaled_value(int32_t is_percent, int32_t value, int32_t total, int32_t round_up) {
  if (!is_percent) return value;
  const double x = (double)value * (double)total / 100;
  return (int32_t)(round_up ? std::ceil(x) : std::floor(x));
}

// calculateCoordinationUpdatedReplicasBound — rolebasedgroup_controller.go:1328-1345
int32_t rbgtopo_updated_replicas_bound(int32_t s_pct, int32_t ref_updated, int32_t ref_desired,
                                       int32_t request_desired, int32_t* lower, int32_t* upper) {
  if (!lower || !upper) return RBGTOPO_EINVAL;
  if (ref_desired == 0) { *lower = 0; *upper = 0; return RBGTOPO_OK; }
  const int64_t a = ref_updated, b = ref_desired, d = request_desired, s = s_pct;
  *lower = (int32_t)std::round((double)std::max<int64_t>(100 * a * d - s * b * d, 0) / (double)(100 * b));
  *upper = (int32_t)std::round((double)std::max<int64_t>(s * b * d + 100 * a * d, 0) / (double)(100 * b));
  return RBGTOPO_OK;
}

// calculateNextRollingTarget — rolebasedgroup_controller.go:1223-1263 with
// getFastestAndSlowestRole :1265-1282.  Ties the Go comparator leaves open
// (sort.Slice is unstable over an unsorted set) are broken by role index.
int32_t rbgtopo_next_rolling_target(int32_t s_pct, int32_t n, const int32_t* desired,
                                    const int32_t* updated, const int32_t* ready,
                                    int32_t* rolling_target) {
  if (n < 0 || (n && (!desired || !updated || !ready || !rolling_target))) return RBGTOPO_EINVAL;
  if (n <= 1) return 1;
  std::vector<double> ratio(n);
  for (int i = 0; i < n; ++i) ratio[i] = (double)updated[i] / (double)desired[i];
  auto less = [&](int a, int b) {
    if (std::fabs(ratio[a] - ratio[b]) > 1e-6) return ratio[a] < ratio[b];
    return desired[a] > desired[b];
  };
  std::vector<int> ord;
  for (int i = 0; i < n; ++i) {  // insertion sort == oracle/refpinned.py
    size_t pos = ord.size();
    while (pos > 0 && less(i, ord[pos - 1])) --pos;
    ord.insert(ord.begin() + pos, i);
  }
  const int fastest = ord.back(), slowest = ord.front();
  for (int i = 0; i < n; ++i) rolling_target[i] = updated[i];
  int32_t max_skew = rbgtopo_scaled_value(1, s_pct, desired[slowest], 1);  // ParseIntStrAsNonZero
  if (max_skew < 1) max_skew = 1;
  int32_t lo, hi;
  rbgtopo_updated_replicas_bound(s_pct, updated[fastest], desired[fastest], desired[slowest], &lo, &hi);
  const int32_t balance = (lo + hi + 1) >> 1;
  const int32_t dist = std::max(balance - updated[slowest], 0);
  int32_t step = std::max(dist, max_skew >> 1);
  if (ready[fastest] == desired[fastest]) step = std::max(step, 1);
  rolling_target[slowest] = std::min(updated[slowest] + step, hi + 1);
  return RBGTOPO_OK;
}

// CalculatePartitionReplicas — pkg/utils/utils.go:139-162.  has_partition = 0: nil partition;
// is_percent: the partition is the string "<value>%"; replicas < 0 stands for a nil replicas pointer (-> 1).
int32_t rbgtopo_partition_replicas(int32_t has_partition, int32_t is_percent, int32_t value, int32_t replicas,
                                   int32_t* out) {
  if (!out) return RBGTOPO_EINVAL;
  if (!has_partition) { *out = 0; return RBGTOPO_OK; }
  const int32_t reps = replicas < 0 ? 1 : replicas;
  // roundUp = true keeps at least one old pod when partition > "0%" and replicas > 0
  int32_t p = rbgtopo_scaled_value(is_percent, value, reps, 1);
  // partition < "100%" and replicas >= 1: at least one pod is upgraded
  if (reps >= 1 && p == reps && is_percent && value != 100) p = reps - 1;
  *out = std::max(std::min(p, reps), 0);
  return RBGTOPO_OK;
}

// ParseIntStrAsNonZero — pkg/utils/utils.go:177-185.
int32_t rbgtopo_intstr_non_zero(int32_t is_percent, int32_t value, int32_t replicas, int32_t* out) {
  if (!out) return RBGTOPO_EINVAL;
  const int32_t v = rbgtopo_scaled_value(is_percent, value, replicas, 1);
  *out = v < 1 ? 1 : v;
  return RBGTOPO_OK;
}

// mergeStrategyRollingUpdate — rolebasedgroup_controller.go:1284-1314, for ONE role present in both
// maps (a role present in one map only is taken as it is).  A strategy is six ints:
// maxUnavailable (has, is_percent, value), partition (has, is_percent, value); both fields are
// compared scaled to 100 with roundUp, a nil field counts as 0 (the Go code drops the error).
int32_t rbgtopo_merge_rolling_update(const int32_t* a, const int32_t* b, int32_t* out) {
  if (!a || !b || !out) return RBGTOPO_EINVAL;
  auto scaled = [](const int32_t* f) { return f[0] ? rbgtopo_scaled_value(f[1], f[2], 100, 1) : 0; };
  for (int i = 0; i < 6; ++i) out[i] = a[i];
  if (scaled(a) > scaled(b))
    for (int i = 0; i < 3; ++i) out[i] = b[i];          // the smaller maxUnavailable wins
  if (scaled(a + 3) < scaled(b + 3))
    for (int i = 3; i < 6; ++i) out[i] = b[i];          // the larger partition wins
  return RBGTOPO_OK;
}

// GetWorkloadName — api/workloads/v1alpha2/helper.go:68-81: "{rbg}-{role}", cut to 63 bytes, trailing '-' trimmed.
// Writes the NUL-terminated name (out_len >= 64); returns its length or RBGTOPO_EINVAL.
int32_t rbgtopo_workload_name(const char* rbg_name, const char* role_name, char* out, int32_t out_len) {
  if (!rbg_name || !role_name || !out || out_len < 64) return RBGTOPO_EINVAL;
  std::string name = std::string(rbg_name) + "-" + role_name;
  if (name.size() > 63) {
    name.resize(63);
    while (!name.empty() && name.back() == '-') name.pop_back();
  }
  memcpy(out, name.c_str(), name.size() + 1);
  return (int32_t)name.size();
}

// GenGroupUniqueKey — helper.go:135-144: lower-case hex SHA-1 of "namespace/name" (40 characters + NUL, out_len >= 41).
int32_t rbgtopo_group_unique_key(const char* ns, const char* name, char* out, int32_t out_len) {
  if (!ns || !name || !out || out_len < 41) return RBGTOPO_EINVAL;
  const std::string msg = std::string(ns) + "/" + name;
  // SHA-1 (FIPS 180-4), one pass over the padded message
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  std::vector<unsigned char> m(msg.begin(), msg.end());
  const uint64_t bits = (uint64_t)m.size() * 8;
  m.push_back(0x80);
  while (m.size() % 64 != 56) m.push_back(0);
  for (int i = 7; i >= 0; --i) m.push_back((unsigned char)(bits >> (8 * i)));
  auto rol = [](uint32_t v, int r) { return (v << r) | (v >> (32 - r)); };
  for (size_t off = 0; off < m.size(); off += 64) {
    uint32_t w[80];
    for (int i = 0; i < 16; ++i)
      w[i] = ((uint32_t)m[off + 4 * i] << 24) | ((uint32_t)m[off + 4 * i + 1] << 16) | ((uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
    for (int i = 16; i < 80; ++i) w[i] = rol(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
    for (int i = 0; i < 80; ++i) {
      uint32_t f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
      else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
      const uint32_t t = rol(a, 5) + f + e + k + w[i];
      e = d; d = c; c = rol(b, 30); b = a; a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
  }
  static const char* hex = "0123456789abcdef";
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 8; ++j) out[8 * i + j] = hex[(h[i] >> (28 - 4 * j)) & 15];
  out[40] = 0;
  return 40;
}

// InheritPodGroupAnnotations — pkg/scheduler/common/annotation_inheritance.go:23-43, per key: 1 when the key
// starts with one of the n_prefixes prefixes (the PodGroup inherits it), else 0.  The map filtering around it is
// the host language's; an empty result is reported as nil by the reference.
int32_t rbgtopo_inherits_annotation(const char* key, int32_t n_prefixes, const char* const* prefixes) {
  if (!key || n_prefixes < 0 || (n_prefixes && !prefixes)) return RBGTOPO_EINVAL;
  for (int i = 0; i < n_prefixes; ++i) {
    if (!prefixes[i]) return RBGTOPO_EINVAL;
    const size_t n = strlen(prefixes[i]);
    if (strncmp(key, prefixes[i], n) == 0) return 1;
  }
  return 0;
}

}  // extern "C"
