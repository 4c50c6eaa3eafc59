// Kernel launch helper: per-launch attributes of the path's kernels (L2 access-policy window, optional PDL).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
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

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    // measured on c2 (64 x 10 s, CUDA graph): 12.30 ms with PDL vs 12.15 ms without -- the persistent kernels
    // own their SM (200+ KB smem), so dependents cannot become resident early; opt-in only (GAM_PDL=1)
    const char* e = std::getenv("GAM_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// L2 residency window for the fp32 residual stream x [rows, d_model]: every layer reads and rewrites it eight times
// (four residual GEMM epilogues, four LayerNorms), and at the benchmark shape it is 49 MB -- it fits the 126 MB L2
// but is evicted by the weight / activation streams in between unless its lines are marked persisting.  The window
// is a per-launch attribute, so it is recorded in CUDA-graph kernel nodes as well.  Set by gam_encode around the
// layer loop; base == nullptr disables it.
struct L2Window {
  void* base = nullptr;
  size_t bytes = 0;
  float hit_ratio = 0.f;
};
inline L2Window& l2_window() {
  static thread_local L2Window w;   // per host thread: handles driven from different threads do not interfere
  return w;
}

// launch `kernel` with the path's launch attributes.  With PDL the kernel may start while its stream predecessor
// drains; the kernel itself calls ptx::pdl_wait() before touching global memory (see ptx.cuh)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  const L2Window& w = l2_window();
  if (w.base != nullptr) {
    at[n].id = cudaLaunchAttributeAccessPolicyWindow;
    at[n].val.accessPolicyWindow.base_ptr = w.base;
    at[n].val.accessPolicyWindow.num_bytes = w.bytes;
    at[n].val.accessPolicyWindow.hitRatio = w.hit_ratio;
    at[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace gam
