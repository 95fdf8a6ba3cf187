// Library-wide state: thread-local error text, launch counter, ABI version.
#include <atomic>
#include <mutex>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/qw.h"
#include "qw_common.cuh"

namespace qw {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int pdl_mode() {
  static const int v = [] {
    const char* e = getenv("QW_PDL");
    return (e && e[0] == '0') ? 0 : (e && e[0] == '2') ? 2 : 1;
  }();
  return v;
}

// ---- option table
struct OptDef {
  const char* name;
  int dflt;
};
static const OptDef kOptDefs[kOptCount] = {{"FAST_PATH", 1}, {"GY_MMA", 1}, {"BWD_FUSED", 0}, {"FWD_ETMA", 1}, {"FIN_EARLY", 1},
                                           {"ADJ_TRIG", 1}, {"PRE_EX", 1}, {"ADJ_SPEC", 1}, {"PRE_CTAS", 4}, {"GY_WARPS", 0}, {"DP_TIMEOUT_MS", 600000}, {"DBG_FWD", 0}, {"DBG_GY", 0}, {"STEM_CHAIN", 1}, {"REV_TILES", 0}};
static std::atomic<int> g_opt[kOptCount];
static std::once_flag g_opt_once;
static void opt_init() {
  for (int i = 0; i < kOptCount; ++i) {
    char env[64];
    snprintf(env, sizeof(env), "QW_%s", kOptDefs[i].name);
    const char* e = getenv(env);
    g_opt[i].store(e ? atoi(e) : kOptDefs[i].dflt, std::memory_order_relaxed);
  }
}
int option(Option o) {
  std::call_once(g_opt_once, opt_init);
  return g_opt[o].load(std::memory_order_relaxed);
}
int set_option(const char* name, int value) {
  std::call_once(g_opt_once, opt_init);
  if (!name) return -1;
  if (strncmp(name, "QW_", 3) == 0) name += 3;
  for (int i = 0; i < kOptCount; ++i)
    if (strcmp(name, kOptDefs[i].name) == 0) {
      g_opt[i].store(value, std::memory_order_relaxed);
      return 0;
    }
  return -1;
}

// ---- debug timeline: caller-owned device buffer of 2 x nslots u64 (start, end) pairs
static std::atomic<unsigned long long*> g_tl_buf{nullptr};
static std::atomic<int> g_tl_slots{0};
static std::atomic<int> g_tl_next{0};
unsigned long long* timeline_next_slot() {
  unsigned long long* b = g_tl_buf.load(std::memory_order_relaxed);
  if (!b) return nullptr;
  const int n = g_tl_slots.load(std::memory_order_relaxed);
  return b + 2 * (size_t)(g_tl_next.fetch_add(1, std::memory_order_relaxed) % n);
}

// ---- per-kernel event timing.  Events are recorded on the launching stream around each kernel and resolved
// lazily in qw_profile_read (which synchronises on them).  Not usable during stream capture.
struct ProfRec {
  cudaEvent_t a, b;
};
static std::mutex g_pmu;
static bool g_prof = false;
static std::vector<ProfRec> g_recs[kKCount];
static cudaEvent_t g_open[kKCount];

static char g_sym[kKCount][96];
void note_symbol(int id, const char* fmt, ...) {
  if (id < 0 || id >= kKCount) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_sym[id], sizeof(g_sym[id]), fmt, ap);
  va_end(ap);
}
const char* kernel_symbol(int id) { return (id >= 0 && id < kKCount) ? g_sym[id] : ""; }
bool profiling_enabled() { return g_prof; }
void profile_begin(int id, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_pmu);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_open[id] = e;
}
void profile_end(int id, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_pmu);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_recs[id].push_back({g_open[id], e});
}
}  // namespace qw

extern "C" {
#ifndef QW_BUILD_STAMP
#define QW_BUILD_STAMP "unstamped"
#endif
const char* qw_build_stamp(void) { return QW_BUILD_STAMP; }
int qw_abi_version(void) { return QW_ABI_VERSION; }
int qw_set_option(const char* name, int value) {
  const int r = qw::set_option(name, value);
  if (r != 0) qw::set_error("qw_set_option: unknown option '%s'", name ? name : "(null)");
  return r;
}
int qw_get_option(const char* name) {
  if (!name) return -1;
  if (strncmp(name, "QW_", 3) == 0) name += 3;
  for (int i = 0; i < qw::kOptCount; ++i)
    if (strcmp(name, qw::kOptDefs[i].name) == 0) return qw::option((qw::Option)i);
  return -1;
}
void qw_set_fast_path(int enable) { qw::set_option("FAST_PATH", enable != 0); }
const char* qw_last_error(void) { return qw::g_err; }
long long qw_launch_count(void) { return qw::g_launches.load(std::memory_order_relaxed); }

int qw_timeline_set(unsigned long long* dev_buf, int nslots) {
  if (dev_buf && nslots <= 0) return -1;
  qw::g_tl_slots.store(nslots > 0 ? nslots : 1);
  qw::g_tl_next.store(0);
  qw::g_tl_buf.store(dev_buf);
  return 0;
}

void qw_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(qw::g_pmu);
  qw::g_prof = on != 0;
}
int qw_profile_read(int kernel_id, double* total_ms, long long* count, int reset) {
  if (kernel_id < 0 || kernel_id >= qw::kKCount || !total_ms || !count) return -1;
  std::lock_guard<std::mutex> lk(qw::g_pmu);
  double tot = 0.0;
  long long n = 0;
  for (auto& r : qw::g_recs[kernel_id]) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      tot += ms;
      ++n;
    }
  }
  *total_ms = tot;
  *count = n;
  if (reset) {
    for (auto& r : qw::g_recs[kernel_id]) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    qw::g_recs[kernel_id].clear();
  }
  return 0;
}
const char* qw_kernel_symbol(int kernel_id) { return qw::kernel_symbol(kernel_id); }
const char* qw_kernel_name(int kernel_id) {
  static const char* names[] = {"qconv_fwd_kernel", "qconv_bwd_post_kernel", "qconv_bwd_pre_kernel", "qconv_bwd_finalize_kernel",
                                "circuit_fwd_kernel", "circuit_bwd_kernel", "circuit_finalize_kernel", "logmel_stft_kernel",
                                "logmel_finish_kernel", "qconv_bwd_adj_kernel", "logmel_prep_kernel", "grads_allreduce_p2p_kernel", "stem2_kernel", "qconv_bwd_fused_kernel"};
  return (kernel_id >= 0 && kernel_id < qw::kKCount) ? names[kernel_id] : "";
}
}
