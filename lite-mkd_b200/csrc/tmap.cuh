// Host-side TMA descriptor construction shared by the tcgen05 kernels.
//
// cuTensorMapEncodeTiled is reached through the runtime's driver entry point (no -lcuda) and memoised on
// every argument: the matching path launches the same few dozen (pointer, shape) combinations every step,
// and an eager (non-graph) step otherwise pays 3-4 encodes per GEMM launch.
#pragma once
#include "common.cuh"

namespace lmkd {

struct TmapSpec {
  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const void* base = nullptr;
  uint64_t dims[4] = {1, 1, 1, 1};        // dims[0] is the contiguous one
  uint64_t strides[3] = {0, 0, 0};        // bytes, for dims 1..3
  uint32_t box[4] = {1, 1, 1, 1};
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
};

// 4-D tiled map, 128-byte swizzle, no interleave, zero OOB fill.  Returns 0 or sets the error text.
int encode_tmap(CUtensorMap* out, const TmapSpec& spec, const char* what);

// LMKD_TMAP_CACHE=0 disables the memoisation (A/B measurements)
void tmap_cache_stats(long long* hits, long long* misses);

}  // namespace lmkd
