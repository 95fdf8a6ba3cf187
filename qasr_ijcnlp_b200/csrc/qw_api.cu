// Library-wide state: thread-local error text, launch counter, ABI version.
#include <atomic>
#include <cstdarg>
#include <cstdio>

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
}  // namespace qw

extern "C" {
int qw_abi_version(void) { return QW_ABI_VERSION; }
const char* qw_last_error(void) { return qw::g_err; }
long long qw_launch_count(void) { return qw::g_launches.load(std::memory_order_relaxed); }
}
