// Shared host-side helpers for libaedit: error reporting, launch accounting, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/aedit.h"

namespace aedit {

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core OPERAND type ("op_t"): the 16-bit type of every GEMM / attention operand (activations after a norm /
// activation, weights, text K/V).  Accumulation, the residual stream, norm statistics and the scheduler state are fp32
// either way, and tcgen05 kind::f16 / mma.sync run fp16 and bf16 at the same rate.  Default: IEEE fp16 — every operand
// of this path is normalised or O(1) (GroupNorm / LayerNorm outputs, softmax weights, fan-in scaled weights), so the
// 5-bit exponent is enough and the 10-bit mantissa cuts the operand rounding error 8x versus bf16 (measured per block
// in profiles/r02_error_attribution_*); conversions saturate to +-65504 instead of producing inf.  -DAE_OPERAND_BF16
// builds the bf16 variant (libaedit_bf16.so, AEDIT_OPERANDS=bf16), the precision BASELINE.json names for configs[1].
// ---------------------------------------------------------------------------------------------------------------------
#ifdef AE_OPERAND_BF16
using op_t = __nv_bfloat16;
using op2_t = __nv_bfloat162;
#define AE_OPERAND_DTYPE 0
#define AE_TMAP_OPERAND_TYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
__host__ __device__ __forceinline__ op_t f2op(float x) { return __float2bfloat16_rn(x); }
__host__ __device__ __forceinline__ float op2f(op_t x) { return __bfloat162float(x); }
__device__ __forceinline__ op2_t ff2op2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float2 op22f2(op2_t v) { return __bfloat1622float2(v); }
#else
using op_t = __half;
using op2_t = __half2;
#define AE_OPERAND_DTYPE 1
#define AE_TMAP_OPERAND_TYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
__host__ __device__ __forceinline__ float ae_sat16(float x) { return fminf(fmaxf(x, -65504.0f), 65504.0f); }
__host__ __device__ __forceinline__ op_t f2op(float x) { return __float2half_rn(ae_sat16(x)); }
__host__ __device__ __forceinline__ float op2f(op_t x) { return __half2float(x); }
__device__ __forceinline__ op2_t ff2op2(float a, float b) { return __floats2half2_rn(ae_sat16(a), ae_sat16(b)); }
__device__ __forceinline__ float2 op22f2(op2_t v) { return __half22float2(v); }
#endif

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
extern thread_local int g_use_pdl;
// Launch priority attached to every kernel launched (and therefore to every kernel NODE captured) while it is set:
// the reverse-process graph is captured with the device's highest priority so that its sub-wave kernels take the next
// free SM slots ahead of the pending CTAs of a forward-process chunk running concurrently on another stream.
extern thread_local int g_launch_priority;       // 0 = leave the stream's priority
// Diagnostic only (ae_set_skip_mask, tools/kernel_share.py): kernel families whose launches are dropped, to measure
// each family's marginal cost inside a captured graph.  Results are garbage while it is non-zero.
// In PDL mode 2 (GEMMs only) the kernel families of this mask are launched with the PDL attribute as well:
// 1 GroupNorm statistics, 2 GroupNorm apply, 4 LayerNorm, 8 attention (ae_set_pdl_extra; A/B in profiles/).
extern thread_local int g_pdl_extra;
extern thread_local int g_skip_mask;             // 1 gemm, 2 split-K reduce, 4 GN stats, 8 GN apply, 16 LayerNorm, 32 attention
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

// 128-byte-swizzled TMA tensor map of operand-type elements (defined in gemm_tcgen05.cu); tm points at a CUtensorMap.
// dims / box innermost first, strides in BYTES for dims 1..rank-1.
int make_operand_tmap(void* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// x * sigmoid(x); __fdividef = one MUFU.RCP + a multiply (2 ulp) instead of the IEEE division sequence: the result is rounded
// to a 16-bit operand right after, and every kernel shares this definition (stream / non-stream GroupNorm: same bits)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float op_round(float x) { return op2f(f2op(x)); }

struct __align__(16) op_x8 {
  op2_t v[4];
};

}  // namespace aedit
