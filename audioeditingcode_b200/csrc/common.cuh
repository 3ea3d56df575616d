// Shared host-side helpers for libaedit: error reporting, launch accounting, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/aedit.h"

namespace aedit {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define AE_CHECK_ARG(cond, ...)                     \
  do {                                              \
    if (!(cond)) return ::aedit::fail(AE_EINVAL, __VA_ARGS__); \
  } while (0)

// call after every kernel launch
inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(AE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return AE_OK;
}

inline cudaStream_t as_stream(ae_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (PDL): every kernel calls pdl_wait() before its first dependent global-memory access
// (blocks until the previous kernel has completed and its writes are visible) and pdl_trigger() once its own main
// work is issued (loads done / last MMA issued), which lets the NEXT kernel of the stream be scheduled so that its
// launch latency and prologue overlap this kernel's tail.  Triggering at the very top instead lets a whole chain of
// kernels become resident and wait on each other, which measured slower (profiles/r01_bench_v8_pdl1.json).
extern int g_use_pdl;
// Launch priority attached to every kernel launched (and therefore to every kernel NODE captured) while it is set:
// the reverse-process graph is captured with the device's highest priority so that its sub-wave kernels take the next
// free SM slots ahead of the pending CTAs of a forward-process chunk running concurrently on another stream.
extern int g_launch_priority;       // 0 = leave the stream's priority
// Diagnostic only (ae_set_skip_mask, tools/kernel_share.py): kernel families whose launches are dropped, to measure
// each family's marginal cost inside a captured graph.  Results are garbage while it is non-zero.
// In PDL mode 2 (GEMMs only) the kernel families of this mask are launched with the PDL attribute as well:
// 1 GroupNorm statistics, 2 GroupNorm apply, 4 LayerNorm, 8 attention (ae_set_pdl_extra; A/B in profiles/).
extern int g_pdl_extra;
extern int g_skip_mask;             // 1 gemm, 2 split-K reduce, 4 GN stats, 8 GN apply, 16 LayerNorm, 32 attention
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// g_use_pdl: 0 = off, 1 = every kernel may start early, 2 = only kernels launched through launch_kernel_early
// (the GEMMs, whose prologue — TMEM allocation, barrier init, descriptor prefetch — is worth overlapping)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_mode(bool early_ok, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                      cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl == 1 || (g_use_pdl == 2 && early_ok)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g_launch_priority != 0) {
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = g_launch_priority;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_early(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args&&... args) {
  return launch_kernel_mode(true, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  return launch_kernel_mode(false, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}

// family: bit of g_pdl_extra that lets this launch start early in PDL mode 2
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_family(int family, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st, Args&&... args) {
  return launch_kernel_mode((g_pdl_extra & family) != 0, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct __align__(16) bf16x8 {
  __nv_bfloat162 v[4];
};

}  // namespace aedit
