// Kernel launch helper: programmatic dependent launch (PDL) on every kernel of the path.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace gam {

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

// launch `kernel` so that it may start while its stream predecessor drains; the kernel itself calls
// ptx::pdl_wait() before touching global memory (see ptx.cuh)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace gam
