// Host-side helpers shared by the C-ABI launchers: status codes, TMA tensor-map encoding.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/devias_b200.h"

namespace dv {

#define DV_CHECK_CUDA(expr)                                                                           \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      dv::set_last_error(#expr, cudaGetErrorString(_e), __FILE__, __LINE__);                          \
      return DEVIAS_ERR_CUDA;                                                                         \
    }                                                                                                 \
  } while (0)

#define DV_REQUIRE(cond, msg)                                                                         \
  do {                                                                                                \
    if (!(cond)) {                                                                                    \
      dv::set_last_error(#cond, msg, __FILE__, __LINE__);                                             \
      return DEVIAS_ERR_ARG;                                                                          \
    }                                                                                                 \
  } while (0)

void set_last_error(const char* what, const char* detail, const char* file, int line);

// 2-D bf16 row-major tensor map with 128-byte swizzle.  inner = contiguous dim.
// Returns 0 on success.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer);
// generic N-d (rank<=5) map; dims/strides innermost first, strides in bytes for dims 1..rank-1
int make_tmap_nd(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

int sm_count();
// Programmatic dependent launch (PDL): every kernel of this library starts with `griddepcontrol.launch_dependents` and executes
// `griddepcontrol.wait` before its first global-memory access, and is launched with the programmatic-stream-serialization
// attribute: the NEXT kernel's launch, block scheduling and prologue (barrier init, tensor-memory allocation, tensor-map
// prefetch) overlap this kernel's tail instead of waiting for the grid to drain.  Semantics are unchanged -- nothing is read or
// written before the wait, which returns only when the preceding grid has completed and flushed.  DEVIAS_PDL=0 turns the
// attribute off (plain stream order).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
void count_launch(int n = 1);
// per-kernel CUDA-event timing, active only between devias_profile_begin/end (kinds: DEVIAS_PROF_*)
int prof_begin(int kind, double work, cudaStream_t s);
void prof_end(int id, cudaStream_t s);

}  // namespace dv
