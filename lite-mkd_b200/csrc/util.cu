// Error plumbing and device queries shared by every translation unit.
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace lmkd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// cudaFuncSetAttribute is per device and the library may be entered from several host threads: remember which
// (kernel, device) pairs already carry their dynamic shared memory limit
int ensure_max_dynamic_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  LMKD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({func, dev})) return 0;
  LMKD_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({func, dev});
  return 0;
}

}  // namespace lmkd
