#include "tmap.cuh"

#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace lmkd {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct Key {
  uint64_t w[13];
  bool operator==(const Key& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = 1469598103934665603ull;   // FNV-1a over the 13 words
    for (uint64_t v : k.w) {
      h ^= v;
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

Key make_key(const TmapSpec& s) {
  Key k;
  k.w[0] = reinterpret_cast<uint64_t>(s.base);
  for (int i = 0; i < 4; ++i) k.w[1 + i] = s.dims[i];
  for (int i = 0; i < 3; ++i) k.w[5 + i] = s.strides[i];
  k.w[8] = (static_cast<uint64_t>(s.box[0]) << 32) | s.box[1];
  k.w[9] = (static_cast<uint64_t>(s.box[2]) << 32) | s.box[3];
  k.w[10] = static_cast<uint64_t>(s.dtype);
  k.w[11] = static_cast<uint64_t>(s.promo);
  int dev = 0;
  cudaGetDevice(&dev);
  k.w[12] = static_cast<uint64_t>(dev);
  return k;
}

std::mutex g_mu;
std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;
long long g_hits = 0, g_misses = 0;
const bool g_cache_on = [] {
  const char* e = getenv("LMKD_TMAP_CACHE");
  return !(e && e[0] == '0');
}();

}  // namespace

int encode_tmap(CUtensorMap* out, const TmapSpec& s, const char* what) {
  EncodeTiledFn enc = get_encode_fn();
  LMKD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  LMKD_CHECK((reinterpret_cast<uintptr_t>(s.base) & 15) == 0, "tensor map %s: base not 16-byte aligned", what);
  for (int i = 0; i < 3; ++i)
    LMKD_CHECK(s.strides[i] % 16 == 0, "tensor map %s: stride %d (%llu bytes) not a multiple of 16", what, i + 1,
               (unsigned long long)s.strides[i]);
  Key key;
  if (g_cache_on) {
    key = make_key(s);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
      *out = it->second;
      ++g_hits;
      return 0;
    }
  }
  cuuint64_t dims[4] = {s.dims[0], s.dims[1], s.dims[2], s.dims[3]};
  cuuint64_t strides[3] = {s.strides[0], s.strides[1], s.strides[2]};
  cuuint32_t box[4] = {s.box[0], s.box[1], s.box[2], s.box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, s.dtype, 4, const_cast<void*>(s.base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, s.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LMKD_CHECK(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled(%s) failed with %d (dims %llu %llu %llu %llu, strides %llu %llu %llu, box %u %u)", what,
             (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
             (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
             (unsigned long long)strides[2], box[0], box[1]);
  if (g_cache_on) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_cache.size() > 8192) g_cache.clear();   // pointers churn (e.g. fresh workspaces every call): stay bounded
    g_cache.emplace(key, *out);
    ++g_misses;
  }
  return 0;
}

void tmap_cache_stats(long long* hits, long long* misses) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (hits) *hits = g_hits;
  if (misses) *misses = g_misses;
}

}  // namespace lmkd
