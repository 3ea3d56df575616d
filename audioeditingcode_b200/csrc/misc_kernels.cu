// Small HBM-bound helper kernels around the tensor-core GEMMs: patch gather (stride-2 / odd-shape convs),
// GEGLU, sinusoidal timestep embedding, nearest resize, layout and dtype movers, row softmax, transpose,
// vocoder activations.  All are coalesced along the channel (innermost) dimension and vectorised where the
// alignment allows it.
#include "common.cuh"

namespace aedit {
namespace {

// ------------------------------------------------------------------ im2col (explicit patch gather)
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ in, int B, int H, int W, int C, int kh, int kw, int stride, int dil,
                              int pad_t, int pad_l, int Ho, int Wo, op_t* __restrict__ out, long long ld_out) {
  pdl_trigger();
  pdl_wait();
  // one CTA row of threads walks the K dimension of one output position (coalesced over channels)
  const long long m = blockIdx.x;
  const int wo = (int)(m % Wo);
  const int ho = (int)((m / Wo) % Ho);
  const int b = (int)(m / ((long long)Wo * Ho));
  const int K = kh * kw * C;
  op_t* orow = out + m * ld_out;
  for (int k = threadIdx.x; k < ld_out; k += blockDim.x) {
    float v = 0.f;
    if (k < K) {
      const int c = k % C;
      const int tap = k / C;
      const int j = tap % kw, i = tap / kw;
      const int h = ho * stride - pad_t + i * dil;
      const int w = wo * stride - pad_l + j * dil;
      if (h >= 0 && h < H && w >= 0 && w < W) v = (float)in[(((long long)b * H + h) * W + w) * C + c];
    }
    orow[k] = f2op(v);
  }
}

// vectorised variant (C % 8 == 0): one item = 8 consecutive channels of one tap of one output position
template <typename T>
__global__ void __launch_bounds__(256) im2col_vec8_kernel(const T* __restrict__ in, int B, int H, int W, int C, int kh,
                                                          int kw, int stride, int dil, int pad_t, int pad_l, int Ho,
                                                          int Wo, op_t* __restrict__ out, long long ld_out,
                                                          long long M) {
  pdl_trigger();
  pdl_wait();
  const int kv = (int)(ld_out >> 3);
  const int K = kh * kw * C;
  const long long total = M * kv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / kv;
    const int k = (int)(i - m * kv) << 3;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (k < K) {
      const int wo = (int)(m % Wo);
      const int ho = (int)((m / Wo) % Ho);
      const int b = (int)(m / ((long long)Wo * Ho));
      const int tap = k / C, c = k - tap * C;
      const int j = tap % kw, ii = tap / kw;
      const int h = ho * stride - pad_t + ii * dil;
      const int w = wo * stride - pad_l + j * dil;
      if (h >= 0 && h < H && w >= 0 && w < W) {
        const T* src = in + (((long long)b * H + h) * W + w) * C + c;
        if constexpr (sizeof(T) == 2) {
          o = *reinterpret_cast<const uint4*>(src);
        } else {
          const float4 f0 = *reinterpret_cast<const float4*>(src);
          const float4 f1 = *reinterpret_cast<const float4*>(src + 4);
          op2_t h0 = ff2op2(f0.x, f0.y), h1 = ff2op2(f0.z, f0.w);
          op2_t h2 = ff2op2(f1.x, f1.y), h3 = ff2op2(f1.z, f1.w);
          o.x = *reinterpret_cast<uint32_t*>(&h0);
          o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2);
          o.w = *reinterpret_cast<uint32_t*>(&h3);
        }
      }
    }
    *reinterpret_cast<uint4*>(out + m * ld_out + k) = o;
  }
}

// ------------------------------------------------------------------ GEGLU
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__global__ void geglu_kernel(const op_t* __restrict__ h, long long rows, int inner,
                             op_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int vec = inner >> 3;
  const long long total = rows * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec;
    const int c = (int)(i % vec) << 3;
    const uint4 xa = *reinterpret_cast<const uint4*>(h + r * 2 * inner + c);
    const uint4 xg = *reinterpret_cast<const uint4*>(h + r * 2 * inner + inner + c);
    const op2_t* pa = reinterpret_cast<const op2_t*>(&xa);
    const op2_t* pg = reinterpret_cast<const op2_t*>(&xg);
    uint4 o;
    op2_t* po = reinterpret_cast<op2_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = op22f2(pa[k]);
      const float2 g = op22f2(pg[k]);
      po[k] = ff2op2(a.x * gelu_erf(g.x), a.y * gelu_erf(g.y));
    }
    *reinterpret_cast<uint4*>(out + r * inner + c) = o;
  }
}

// ------------------------------------------------------------------ timestep embedding  [cos | sin]
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int B, int dim,
                                          op_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  // util.py:184-190: freqs = exp(-ln(10000) * k / half) in fp32; args = t.float() * freqs
  const float freq = expf(-9.210340371976184f * (float)k / (float)half);
  const float arg = (float)t[b] * freq;
  out[(long long)b * dim + k] = f2op(cosf(arg));
  out[(long long)b * dim + half + k] = f2op(sinf(arg));
}

// ------------------------------------------------------------------ nearest resize (f32 NHWC -> bf16 NHWC)
__global__ void upsample_nearest_kernel(const float* __restrict__ x, int B, int H, int W, int C, int Ho, int Wo,
                                        op_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int vec = C >> 2;
  const long long total = (long long)B * Ho * Wo * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vec) << 2;
    long long r = i / vec;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    // torch 'nearest': src = floor(dst * in / out)
    const int hs = min((int)(((long long)ho * H) / Ho), H - 1);
    const int ws = min((int)(((long long)wo * W) / Wo), W - 1);
    const float4 v = *reinterpret_cast<const float4*>(x + (((long long)b * H + hs) * W + ws) * C + c);
    op2_t h0 = ff2op2(v.x, v.y);
    op2_t h1 = ff2op2(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out + (((long long)b * Ho + ho) * Wo + wo) * C + c) = pk;
  }
}

// ------------------------------------------------------------------ NCHW <-> NHWC (tiled transpose per sample)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, long long HW, float* __restrict__ of,
                                    op_t* __restrict__ ob) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? x[((long long)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      const float v = tile[threadIdx.x][i];
      const long long o = ((long long)b * HW + p) * C + c;
      if (of) of[o] = v;
      if (ob) ob[o] = f2op(v);
    }
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int C, long long HW, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? x[((long long)b * HW + p) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    if (c < C && p < HW) out[((long long)b * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ elementwise
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, long long n, op_t* __restrict__ out, int silu) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (silu) v = silu_f(v);
    out[i] = f2op(v);
  }
}
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float sb, long long n,
                               float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] + sb * b[i];
}
__global__ void leaky_relu_bf16_kernel(const float* __restrict__ x, long long n, float scale, float slope,
                                       op_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i] * scale;
    out[i] = f2op(v > 0.f ? v : v * slope);
  }
}
// (wavs * 32768).astype("int16") of hifigan/utilities.py:80: truncation toward zero; tanh output is inside (-1, 1), the
// clamp only guards the +-1.0 corner the fp32 tanh can round to
__global__ void wave_to_int16_kernel(const float* __restrict__ x, long long n, int16_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(x[i] * 32768.0f, -32768.0f), 32767.0f);
    out[i] = (int16_t)v;
  }
}

__global__ void tanh_f32_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = tanhf(x[i]);
}

// ------------------------------------------------------------------ row softmax (fp32 -> bf16), one CTA per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, int n, long long ld,
                                                           op_t* __restrict__ out, long long ld_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  const float* xr = x + (long long)blockIdx.x * ld;
  op_t* orow = out + (long long)blockIdx.x * ld_out;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, xr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float su = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) su += __expf(xr[i] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) su += __shfl_xor_sync(0xffffffffu, su, o);
  if (lane == 0) red[warp] = su;
  __syncthreads();
  su = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) su += red[w];
  const float inv = 1.0f / su;
  for (int i = threadIdx.x; i < n; i += 256) orow[i] = f2op(__expf(xr[i] - mx) * inv);
}

__global__ void transpose_bf16_kernel(const op_t* __restrict__ x, int rows, int cols,
                                      op_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ op_t tile[32][34];
  const long long boff = (long long)blockIdx.z * rows * cols;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? x[boff + (long long)r * cols + c] : f2op(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[boff + (long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

inline int ew_grid(long long n, int threads) {
  long long b = ceil_div64(n, threads);
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace aedit

using namespace aedit;

extern "C" int ae_im2col(const void* in, int in_is_bf16, int B, int H, int W, int C, int kh, int kw, int stride, int dil,
                         int pad_t, int pad_l, int Ho, int Wo, void* out_bf16, int64_t ld_out, ae_stream stream) {
  AE_CHECK_ARG(in && out_bf16 && B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && dil > 0,
               "ae_im2col: bad argument");
  AE_CHECK_ARG(ld_out >= (int64_t)kh * kw * C, "ae_im2col: ld_out too small");
  const long long M = (long long)B * Ho * Wo;
  AE_CHECK_ARG(M > 0 && M < 2147483647LL, "ae_im2col: bad output size");
  if (C % 8 == 0 && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0) {
    const long long items = M * (ld_out / 8);
    long long blocks = ceil_div64(items, 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    cudaError_t e;
    if (in_is_bf16)
      e = launch_kernel(im2col_vec8_kernel<op_t>, dim3((unsigned)blocks), dim3(256), (size_t)0, as_stream(stream),
                        reinterpret_cast<const op_t*>(in), B, H, W, C, kh, kw, stride, dil, pad_t, pad_l, Ho, Wo,
                        reinterpret_cast<op_t*>(out_bf16), (long long)ld_out, M);
    else
      e = launch_kernel(im2col_vec8_kernel<float>, dim3((unsigned)blocks), dim3(256), (size_t)0, as_stream(stream),
                        reinterpret_cast<const float*>(in), B, H, W, C, kh, kw, stride, dil, pad_t, pad_l, Ho, Wo,
                        reinterpret_cast<op_t*>(out_bf16), (long long)ld_out, M);
    if (e != cudaSuccess) return fail(AE_ECUDA, "ae_im2col launch: %s", cudaGetErrorString(e));
    return launched("ae_im2col");
  }
  const int threads = ld_out >= 256 ? 256 : 128;
  if (in_is_bf16)
    launch_kernel(im2col_kernel<op_t>, dim3((unsigned)M), dim3(threads), (size_t)(0), as_stream(stream), reinterpret_cast<const op_t*>(in), B, H, W, C, kh, kw, stride, dil, pad_t, pad_l, Ho, Wo,
        reinterpret_cast<op_t*>(out_bf16), ld_out);
  else
    launch_kernel(im2col_kernel<float>, dim3((unsigned)M), dim3(threads), (size_t)(0), as_stream(stream), reinterpret_cast<const float*>(in), B, H, W, C, kh, kw, stride, dil, pad_t, pad_l, Ho, Wo,
        reinterpret_cast<op_t*>(out_bf16), ld_out);
  return launched("ae_im2col");
}

extern "C" int ae_geglu(const void* h, int64_t rows, int inner, void* out, ae_stream stream) {
  AE_CHECK_ARG(h && out && rows > 0 && inner > 0 && inner % 8 == 0, "ae_geglu: bad argument (inner %% 8 == 0 required)");
  launch_kernel(geglu_kernel, dim3(ew_grid(rows * (inner / 8), 256)), dim3(256), (size_t)(0), as_stream(stream), reinterpret_cast<const op_t*>(h), rows, inner, reinterpret_cast<op_t*>(out));
  return launched("ae_geglu");
}

extern "C" int ae_timestep_embedding(const int64_t* t, int B, int dim, void* out_bf16, ae_stream stream) {
  AE_CHECK_ARG(t && out_bf16 && B > 0 && dim > 0 && dim % 2 == 0, "ae_timestep_embedding: bad argument");
  const int n = B * (dim / 2);
  launch_kernel(timestep_embedding_kernel, dim3((n + 127) / 128), dim3(128), (size_t)(0), as_stream(stream), reinterpret_cast<const long long*>(t), B, dim, reinterpret_cast<op_t*>(out_bf16));
  return launched("ae_timestep_embedding");
}

extern "C" int ae_upsample_nearest(const float* x, int B, int H, int W, int C, int Ho, int Wo, void* out_bf16,
                                   ae_stream stream) {
  AE_CHECK_ARG(x && out_bf16 && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && Ho > 0 && Wo > 0,
               "ae_upsample_nearest: bad argument");
  launch_kernel(upsample_nearest_kernel, dim3(ew_grid((long long)B * Ho * Wo * (C / 4), 256)), dim3(256), (size_t)(0), as_stream(stream), x, B, H, W, C, Ho, Wo, reinterpret_cast<op_t*>(out_bf16));
  return launched("ae_upsample_nearest");
}

extern "C" int ae_nchw_to_nhwc(const float* x, int B, int C, int H, int W, float* out_f32, void* out_bf16,
                               ae_stream stream) {
  AE_CHECK_ARG(x && (out_f32 || out_bf16) && B > 0 && C > 0 && H > 0 && W > 0, "ae_nchw_to_nhwc: bad argument");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)ceil_div64(HW, 32), (C + 31) / 32, B);
  launch_kernel(nchw_to_nhwc_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), as_stream(stream), x, C, HW, out_f32,
                                                                   reinterpret_cast<op_t*>(out_bf16));
  return launched("ae_nchw_to_nhwc");
}

extern "C" int ae_nhwc_to_nchw(const float* x, int B, int C, int H, int W, float* out_f32, ae_stream stream) {
  AE_CHECK_ARG(x && out_f32 && B > 0 && C > 0 && H > 0 && W > 0, "ae_nhwc_to_nchw: bad argument");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)ceil_div64(HW, 32), (C + 31) / 32, B);
  launch_kernel(nhwc_to_nchw_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), as_stream(stream), x, C, HW, out_f32);
  return launched("ae_nhwc_to_nchw");
}

extern "C" int ae_cast_f32_bf16(const float* x, int64_t n, void* out_bf16, int silu, ae_stream stream) {
  AE_CHECK_ARG(x && out_bf16 && n > 0, "ae_cast_f32_bf16: bad argument");
  launch_kernel(cast_f32_bf16_kernel, dim3(ew_grid(n, 256)), dim3(256), (size_t)(0), as_stream(stream), x, n, reinterpret_cast<op_t*>(out_bf16),
                                                                       silu);
  return launched("ae_cast_f32_bf16");
}

extern "C" int ae_add_f32(const float* a, const float* b, float scale_b, int64_t n, float* out, ae_stream stream) {
  AE_CHECK_ARG(a && b && out && n > 0, "ae_add_f32: bad argument");
  launch_kernel(add_f32_kernel, dim3(ew_grid(n, 256)), dim3(256), (size_t)(0), as_stream(stream), a, b, scale_b, n, out);
  return launched("ae_add_f32");
}

extern "C" int ae_leaky_relu_bf16(const float* x, int64_t n, float scale, float slope, void* out_bf16, ae_stream stream) {
  AE_CHECK_ARG(x && out_bf16 && n > 0, "ae_leaky_relu_bf16: bad argument");
  launch_kernel(leaky_relu_bf16_kernel, dim3(ew_grid(n, 256)), dim3(256), (size_t)(0), as_stream(stream), x, n, scale, slope,
                                                                         reinterpret_cast<op_t*>(out_bf16));
  return launched("ae_leaky_relu_bf16");
}

extern "C" int ae_wave_to_int16(const float* x, int64_t n, int16_t* out, ae_stream stream) {
  AE_CHECK_ARG(x && out && n > 0, "ae_wave_to_int16: bad argument");
  launch_kernel(wave_to_int16_kernel, dim3(ew_grid(n, 256)), dim3(256), (size_t)(0), as_stream(stream), x, n, out);
  return launched("ae_wave_to_int16");
}

extern "C" int ae_tanh_f32(const float* x, int64_t n, float* out, ae_stream stream) {
  AE_CHECK_ARG(x && out && n > 0, "ae_tanh_f32: bad argument");
  launch_kernel(tanh_f32_kernel, dim3(ew_grid(n, 256)), dim3(256), (size_t)(0), as_stream(stream), x, n, out);
  return launched("ae_tanh_f32");
}

extern "C" int ae_softmax_rows(const float* x, int64_t rows, int n, int64_t ld, void* out_bf16, int64_t ld_out,
                               ae_stream stream) {
  AE_CHECK_ARG(x && out_bf16 && rows > 0 && n > 0 && rows < 2147483647LL, "ae_softmax_rows: bad argument");
  launch_kernel(softmax_rows_kernel, dim3((unsigned)rows), dim3(256), (size_t)(0), as_stream(stream), x, n, ld, reinterpret_cast<op_t*>(out_bf16),
                                                                     ld_out);
  return launched("ae_softmax_rows");
}

extern "C" int ae_transpose_bf16(const void* x, int batch, int rows, int cols, void* out, ae_stream stream) {
  AE_CHECK_ARG(x && out && batch > 0 && rows > 0 && cols > 0, "ae_transpose_bf16: bad argument");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  launch_kernel(transpose_bf16_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), as_stream(stream), reinterpret_cast<const op_t*>(x), rows, cols,
                                                                     reinterpret_cast<op_t*>(out));
  return launched("ae_transpose_bf16");
}
