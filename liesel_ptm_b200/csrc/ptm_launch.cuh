// ptm_launch.cuh -- host-side launch helpers shared by the sweeps.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

namespace ptm {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device), and again only
// when a larger size is requested than was granted before -- not on every launch.
template <typename K>
inline cudaError_t ensure_smem(K kern, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const auto key = std::make_pair(reinterpret_cast<const void*>(kern), dev);
  std::lock_guard<std::mutex> lk(mu);
  auto it = granted.find(key);
  if (it != granted.end() && bytes <= it->second) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) granted[key] = bytes;
  return e;
}

}  // namespace ptm
