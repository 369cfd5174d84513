// Error plumbing and device queries shared by every translation unit.
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

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

// ---- per-category launch timing ---------------------------------------------------------------------
namespace {
struct TimedLaunch {
  cudaEvent_t beg, end;
  double work;
};
std::atomic<bool> g_timing{false};
std::mutex g_timed_mu;
std::vector<TimedLaunch> g_timed[TIME_NCAT];   // measurement hook: guarded, but meant for one host thread
}  // namespace

void kernel_timing_enable(int on) { g_timing.store(on != 0); }
bool kernel_timing_on() { return g_timing.load(); }

int KernelTimingScope::begin() {
  if (!g_timing.load()) return 0;
  LMKD_CUDA(cudaEventCreate(&beg_));
  LMKD_CUDA(cudaEventCreate(&end_));
  LMKD_CUDA(cudaEventRecord(beg_, st_));
  return 0;
}

int KernelTimingScope::end() {
  if (beg_ == nullptr) return 0;
  LMKD_CUDA(cudaEventRecord(end_, st_));
  std::lock_guard<std::mutex> lock(g_timed_mu);
  g_timed[cat_].push_back(TimedLaunch{beg_, end_, work_});
  return 0;
}

int kernel_timing_read(int category, double* ms, double* work, int* launches) {
  LMKD_CHECK(category >= 0 && category < TIME_NCAT, "kernel_timing_read: unknown category %d", category);
  double t = 0, f = 0;
  std::lock_guard<std::mutex> lock(g_timed_mu);
  cudaError_t err = cudaSuccess;
  for (auto& tl : g_timed[category]) {
    float e = 0;
    if (err == cudaSuccess) err = cudaEventSynchronize(tl.end);
    if (err == cudaSuccess) err = cudaEventElapsedTime(&e, tl.beg, tl.end);
    t += e;
    f += tl.work;
    cudaEventDestroy(tl.beg);       // the record is always released, also on the error path
    cudaEventDestroy(tl.end);
  }
  *ms = t;
  *work = f;
  *launches = static_cast<int>(g_timed[category].size());
  g_timed[category].clear();
  if (err != cudaSuccess) {
    set_error("kernel_timing_read: %s", cudaGetErrorString(err));
    return 2;
  }
  return 0;
}

}  // namespace lmkd
