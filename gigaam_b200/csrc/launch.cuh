// Kernel launch helpers shared by the translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <utility>

namespace gam {

// cudaFuncSetAttribute (dynamic shared memory opt-in, cluster sizes) is per DEVICE: a process that drives several GPUs
// must repeat it on each.  `first()` is true exactly once per device for the call site that owns the object.
struct PerDeviceOnce {
  std::atomic<uint64_t> seen{0};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    return (seen.fetch_or(bit) & bit) == 0;
  }
};

// cudaLaunchKernelEx with no attributes: kernels carrying __cluster_dims__ need the extended launch API.
// (Programmatic dependent launch and a persisting-L2 window on the residual stream were measured in round 1 and
// lost: persistent CTAs own their SM, and the carve-out cost the weight-streaming GEMMs more than it saved.)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace gam
