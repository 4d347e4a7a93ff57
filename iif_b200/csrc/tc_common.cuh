// Host-side helpers shared by the tensor-core launch paths (gemm_tc.cu, head_fused.cu): the driver entry point
// of cuTensorMapEncodeTiled and a cache of encoded tensor maps.
#pragma once
#include <cuda.h>

#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace iif {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t inner, outer, ld; uint32_t box_inner, box_outer, f32;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer && f32 == o.f32;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    for (uint64_t v : {k.inner, k.outer, k.ld, (uint64_t)k.box_inner, (uint64_t)k.box_outer, (uint64_t)k.f32})
      h = h * 1000003u ^ (size_t)v;
    return h;
  }
};

// Row-major [outer, inner] tensor (bf16 or fp32) with leading dimension ld (elements); box =
// box_inner x box_outer with a 128-byte inner extent, 128B swizzle.  Loads: out-of-bounds elements
// read as zero (tile tails need no host padding); stores: out-of-bounds elements are not written.
inline int make_map(CUtensorMap* out, const void* ptr, bool f32, uint64_t inner, uint64_t outer, uint64_t ld,
                    uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, inner, outer, ld, box_inner, box_outer, f32 ? 1u : 0u};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return IIF_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return IIF_EDRIVER;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {ld * (f32 ? 4u : 2u)};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return IIF_EDRIVER;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return IIF_OK;
}


// Launch with the COOPERATIVE attribute: the driver either makes every CTA of the grid resident at once (whatever
// else runs on the device) or fails the launch -- the property every in-kernel inter-CTA wait of this library relies
// on.  Programmatic stream serialization (the next kernel's prologue overlapping this one's tail) is requested
// alongside while the driver accepts the combination; the first rejection drops it for the rest of the process.
inline cudaError_t launch_cooperative(cudaLaunchConfig_t& cfg, const void* fn, void** kargs, bool want_pdl) {
  static std::atomic<int> pdl_ok{-1};
  int ok = pdl_ok.load(std::memory_order_relaxed);
  if (ok < 0) {
    const char* e = getenv("IIF_B200_COOP_PDL");
    ok = (e && e[0] == '0') ? 0 : 1;
    pdl_ok.store(ok, std::memory_order_relaxed);
  }
  // IIF_B200_COOP=0 (experiments only): plain launch -- co-residency is then the caller's problem
  static const bool coop = [] { const char* e = getenv("IIF_B200_COOP"); return !(e && e[0] == '0'); }();
  cudaError_t e = cudaSuccess;
  for (int attempt = 0; attempt < 2; ++attempt) {
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (coop) {
      attrs[na].id = cudaLaunchAttributeCooperative;
      attrs[na].val.cooperative = 1;
      ++na;
    }
    const bool pdl = want_pdl && pdl_ok.load(std::memory_order_relaxed) == 1;
    if (pdl) {
      attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attrs[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    e = cudaLaunchKernelExC(&cfg, fn, kargs);
    if (e == cudaSuccess || !pdl || !coop) break;
    cudaGetLastError();                      // rejected together: keep the co-residency guarantee, drop the overlap
    pdl_ok.store(0, std::memory_order_relaxed);
  }
  return e;
}

}  // namespace iif
