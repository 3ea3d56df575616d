// GroupNorm(+SiLU) and LayerNorm producers of the tensor-core operands (K3 of SURVEY.md §2.2).
//   ae_groupnorm : diffusers ResnetBlock2D.norm1/norm2, Transformer2DModel.norm, conv_norm_out
//                  (in-tree: openaimodel.py:213-216,238-239 with util.py:240-242; attention.py:75-78)
//   ae_layernorm : BasicTransformerBlock.norm1/2/3 (attention.py:393-395)
// Both read the fp32 residual stream (channels-last) and write the bf16 operand the next GEMM loads by TMA.
// Statistics are accumulated per sample in a fixed order (no atomics on data) so a sample's result does not
// depend on what else is in the batch.
#include "common.cuh"

namespace aedit {
namespace {

constexpr int kGNThreads = 256;
constexpr int kMaxSplits = 64;

struct GNArgs {
  const float* x1;
  const float* x2;
  int C1, C2, C, G, cpg;
  int B;
  long long HW;
  int S;            // position splits per sample
  long long chunk;  // positions per split
  float eps;
  const float* gamma;
  const float* beta;
  int silu;
  op_t* out;
  op_t* raw_out;
  float* cat_out;
  double* partial;     // [B, S, G, 2]
  float* stats;        // [B, G, 2] mean, rstd
  unsigned int* counters;  // [B]
  // statistics handed over by the producing GEMMs (ae_gemm_args.colstats): fixed-point per-(sample, channel) sums of
  // x1 / x2; when set, no statistics kernel runs and the apply kernels derive mean / rstd themselves
  const long long* cs1;    // [B, C1, 2]
  const long long* cs2;    // [B, C2, 2] (or null when C2 == 0)
};

// mean / rstd of every group of sample b from the producers' fixed-point column sums -> s_mean / s_rstd (shared).
// Integer adds are exact, so the result does not depend on the order of the lanes / of the producing CTAs.
__device__ __forceinline__ void gn_stats_from_colsums(const GNArgs& a, int b, float* s_mean, float* s_rstd) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int g = wid; g < a.G; g += nw) {
    long long su = 0, sq = 0;
    for (int e = lane; e < a.cpg; e += 32) {
      const int c = g * a.cpg + e;
      const longlong2 v = c < a.C1 ? __ldcg(reinterpret_cast<const longlong2*>(a.cs1 + ((long long)b * a.C1 + c) * 2))
                                   : __ldcg(reinterpret_cast<const longlong2*>(a.cs2 + ((long long)b * a.C2 + (c - a.C1)) * 2));
      su += v.x;
      sq += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      su += __shfl_down_sync(0xffffffffu, su, o);
      sq += __shfl_down_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
      const double inv_n = 1.0 / ((double)a.HW * a.cpg);
      const double mean = (double)su * (1.0 / 268435456.0) * inv_n;
      double var = (double)sq * (1.0 / 16777216.0) * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)mean;
      s_rstd[g] = rsqrtf((float)var + a.eps);
    }
  }
}

__device__ __forceinline__ float load_cat(const GNArgs& a, long long row, int c) {
  return c < a.C1 ? a.x1[row * a.C1 + c] : a.x2[row * a.C2 + (c - a.C1)];
}

// grid (S, B), block (TX, TY): thread (tx, ty) owns the channel quads tx, tx+TX, ... (<= 4 of them) and walks the
// positions p0+ty, p0+ty+TY, ... of this CTA's chunk with independent float4 loads (coalesced along channels,
// TY loads in flight per quad).  Per-channel partials are combined over ty in shared memory in a fixed order, then
// per group in double; the last CTA of a sample (arrival counter) finalises mean / rstd for all groups.
constexpr int kGNMaxQuadsPerThread = 4;
constexpr int kGNUnroll = 4;
__global__ void __launch_bounds__(512) gn_stats_kernel(GNArgs a) {
  pdl_wait();
  extern __shared__ float sm[];  // [TY][2*C] per-channel sum / sumsq per ty
  const int s = blockIdx.x, b = blockIdx.y;
  const int TX = blockDim.x, TY = blockDim.y;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TX + tx, nthr = TX * TY;
  const long long p0 = (long long)s * a.chunk;
  const long long p1 = min(a.HW, p0 + a.chunk);
  const int nq = a.C >> 2;
  float su[kGNMaxQuadsPerThread][4], sq[kGNMaxQuadsPerThread][4];
#pragma unroll
  for (int k = 0; k < kGNMaxQuadsPerThread; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) su[k][e] = sq[k][e] = 0.f;
  // positions are walked kGNUnroll at a time: all loads of a batch are issued before any is consumed (an in-order
  // warp would otherwise serialise one DRAM/L2 latency per position)
  for (long long pb = p0 + ty; pb < p1; pb += (long long)TY * kGNUnroll) {
    float4 v[kGNUnroll][kGNMaxQuadsPerThread];
#pragma unroll
    for (int u = 0; u < kGNUnroll; ++u) {
      const long long p = pb + (long long)u * TY;
      const long long row = (long long)b * a.HW + p;
#pragma unroll
      for (int k = 0; k < kGNMaxQuadsPerThread; ++k) {
        const int qd = tx + k * TX;
        if (p < p1 && qd < nq) {
          const int c = qd << 2;
          v[u][k] = c < a.C1 ? *reinterpret_cast<const float4*>(a.x1 + row * a.C1 + c)
                             : *reinterpret_cast<const float4*>(a.x2 + row * a.C2 + (c - a.C1));
        } else {
          v[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kGNUnroll; ++u) {
#pragma unroll
      for (int k = 0; k < kGNMaxQuadsPerThread; ++k) {
        su[k][0] += v[u][k].x; sq[k][0] += v[u][k].x * v[u][k].x;
        su[k][1] += v[u][k].y; sq[k][1] += v[u][k].y * v[u][k].y;
        su[k][2] += v[u][k].z; sq[k][2] += v[u][k].z * v[u][k].z;
        su[k][3] += v[u][k].w; sq[k][3] += v[u][k].w * v[u][k].w;
      }
    }
  }
  pdl_trigger();   // all loads of this CTA are done: overlap the next launch with the reduction tail
  float* mysum = sm + (size_t)ty * 2 * a.C;
#pragma unroll
  for (int k = 0; k < kGNMaxQuadsPerThread; ++k) {
    const int qd = tx + k * TX;
    if (qd < nq) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mysum[(qd << 2) + e] = su[k][e];
        mysum[a.C + (qd << 2) + e] = sq[k][e];
      }
    }
  }
  __syncthreads();
  {
    // one warp per group: lanes stride over the TY x cpg per-channel partials of the group (fixed assignment),
    // fixed-shape shuffle tree -> deterministic
    const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
    const int per = TY * a.cpg;
    for (int g = wid; g < a.G; g += nw) {
      double dsu = 0.0, dsq = 0.0;
      for (int e = lane; e < per; e += 32) {
        const int y = e / a.cpg, c = g * a.cpg + (e - y * a.cpg);
        const float* r = sm + (size_t)y * 2 * a.C;
        dsu += (double)r[c];
        dsq += (double)r[a.C + c];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dsu += __shfl_down_sync(0xffffffffu, dsu, o);
        dsq += __shfl_down_sync(0xffffffffu, dsq, o);
      }
      if (lane == 0) {
        double* dst = a.partial + (((long long)b * a.S + s) * a.G + g) * 2;
        __stcg(dst, dsu);
        __stcg(dst + 1, dsq);
      }
    }
  }
  // ---- the last CTA of the sample to arrive (ticket counter, no waiting) turns the S partials into mean / rstd:
  // one warp per group, lane l sums partials l and l+32 (S <= 64), fixed-shape shuffle tree -> the bits do not depend
  // on which CTA happens to be last.  The apply kernel then needs a single round of independent loads.
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.counters + b, 1u) == (unsigned)a.S - 1u) ? 1 : 0;
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
    // two groups per warp and pass: the (up to) 8 partial loads of a pass are issued together
    for (int g0 = wid; g0 < a.G; g0 += 2 * nw) {
      const long long st = (long long)a.G * 2;
      double pv[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g = g0 + u * nw;
        const double* src = a.partial + ((long long)b * a.S * a.G + g) * 2;
        const bool ok = g < a.G;
        pv[u][0] = (ok && lane < a.S) ? __ldcg(src + lane * st) : 0.0;
        pv[u][1] = (ok && lane < a.S) ? __ldcg(src + lane * st + 1) : 0.0;
        pv[u][2] = (ok && lane + 32 < a.S) ? __ldcg(src + (lane + 32) * st) : 0.0;
        pv[u][3] = (ok && lane + 32 < a.S) ? __ldcg(src + (lane + 32) * st + 1) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g = g0 + u * nw;
        double dsu = pv[u][0] + pv[u][2], dsq = pv[u][1] + pv[u][3];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          dsu += __shfl_down_sync(0xffffffffu, dsu, o);
          dsq += __shfl_down_sync(0xffffffffu, dsq, o);
        }
        if (lane == 0 && g < a.G) {
          const double inv_n = 1.0 / ((double)a.HW * a.cpg);
          const double mean = dsu * inv_n;
          double var = dsq * inv_n - mean * mean;
          if (var < 0.0) var = 0.0;
          float* dst = a.stats + ((long long)b * a.G + g) * 2;
          dst[0] = (float)mean;
          dst[1] = rsqrtf((float)var + a.eps);
        }
      }
    }
    if (tid == 0) a.counters[b] = 0u;   // re-arm for the next launch (stream-ordered after this kernel)
  }
}

// grid (ceil(HW*C/4 / (256*kGNItems)), B): each thread normalises kGNItems float4 items (coalesced along channels).
// Order of work in a CTA: (1) issue the data loads, (2) load mean / rstd (finalised by the last statistics CTA) and
// gamma / beta — independent of (1), so one memory round trip covers both —, (3) build the per-channel affine table
// y = x*A[c] + B[c] in shared memory, (4) apply + SiLU + bf16 store.
constexpr int kGNItems = 8;
constexpr int kGNTab = 10;   // channel-table entries per thread: C <= 256 * kGNTab (checked on the host)
__global__ void __launch_bounds__(kGNThreads) gn_apply_kernel(GNArgs a) {
  // gamma / beta are weights (never written by the previous kernel): fetched before the dependency wait
  float gpre[kGNTab], bpre[kGNTab];
#pragma unroll
  for (int k = 0; k < kGNTab; ++k) {
    const int c = threadIdx.x + k * kGNThreads;
    if (c < a.C) {
      gpre[k] = __ldg(a.gamma + c);
      bpre[k] = __ldg(a.beta + c);
    }
  }
  pdl_wait();
  extern __shared__ float sm[];  // mean[G], rstd[G], A[C], B[C]
  float* s_mean = sm;
  float* s_rstd = sm + a.G;
  float* s_A = sm + 2 * a.G;
  float* s_B = s_A + a.C;
  const int b = blockIdx.y;
  const int vec_per_row = a.C >> 2;
  const int total = (int)(a.HW * vec_per_row);   // < 2^31, checked on the host
  const int base = blockIdx.x * (kGNThreads * kGNItems) + threadIdx.x;
  float4 v[kGNItems];
#pragma unroll
  for (int it = 0; it < kGNItems; ++it) {
    const int idx = base + it * kGNThreads;
    if (idx < total) {
      const int p = idx / vec_per_row;
      const int c = (idx - p * vec_per_row) << 2;
      const long long row = (long long)b * a.HW + p;
      v[it] = c < a.C1 ? *reinterpret_cast<const float4*>(a.x1 + row * a.C1 + c)
                       : *reinterpret_cast<const float4*>(a.x2 + row * a.C2 + (c - a.C1));
    }
  }
  pdl_trigger();   // data loads are in flight
  if (a.cs1) {
    gn_stats_from_colsums(a, b, s_mean, s_rstd);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGNTab; ++k) {
      const int c = threadIdx.x + k * kGNThreads;
      if (c < a.C) {
        const int g = c / a.cpg;
        const float A = s_rstd[g] * gpre[k];
        s_A[c] = A;
        s_B[c] = bpre[k] - s_mean[g] * A;
      }
    }
  } else
  // mean / rstd were finalised by the statistics kernel.  All table loads are issued before any is consumed (a
  // rolled loop would expose one L2 round trip per 256 channels); they are in flight together with the data loads.
  {
    float2 mr[kGNTab];
#pragma unroll
    for (int k = 0; k < kGNTab; ++k) {
      const int c = threadIdx.x + k * kGNThreads;
      if (c < a.C) mr[k] = *reinterpret_cast<const float2*>(a.stats + ((long long)b * a.G + c / a.cpg) * 2);
    }
#pragma unroll
    for (int k = 0; k < kGNTab; ++k) {
      const int c = threadIdx.x + k * kGNThreads;
      if (c < a.C) {
        const float A = mr[k].y * gpre[k];
        s_A[c] = A;
        s_B[c] = bpre[k] - mr[k].x * A;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kGNItems; ++it) {
    const int idx = base + it * kGNThreads;
    if (idx >= total) continue;
    const int p = idx / vec_per_row;
    const int c = (idx - p * vec_per_row) << 2;
    const long long row = (long long)b * a.HW + p;
    const float in[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
    const float4 A4 = *reinterpret_cast<const float4*>(s_A + c);
    const float4 B4 = *reinterpret_cast<const float4*>(s_B + c);
    float o[4] = {fmaf(in[0], A4.x, B4.x), fmaf(in[1], A4.y, B4.y), fmaf(in[2], A4.z, B4.z), fmaf(in[3], A4.w, B4.w)};
    if (a.silu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = silu_f(o[k]);
    }
    const long long off = row * a.C + c;
    op2_t h0 = ff2op2(o[0], o[1]);
    op2_t h1 = ff2op2(o[2], o[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(a.out + off) = pk;
    if (a.raw_out) {
      op2_t r0 = ff2op2(in[0], in[1]);
      op2_t r1 = ff2op2(in[2], in[3]);
      uint2 rk;
      rk.x = *reinterpret_cast<uint32_t*>(&r0);
      rk.y = *reinterpret_cast<uint32_t*>(&r1);
      *reinterpret_cast<uint2*>(a.raw_out + off) = rk;
    }
    if (a.cat_out) *reinterpret_cast<float4*>(a.cat_out + off) = v[it];
  }
}

// Streaming variant of gn_apply_kernel for large tensors: grid (row chunks, B); a CTA builds the per-channel affine
// table once, then walks its rows with kGNStreamItems independent float4 loads in flight per thread.  Same
// arithmetic (A = rstd*gamma, B = beta - mean*A, y = fma(x, A, B), SiLU) -> same bits as gn_apply_kernel.
constexpr int kGNStreamItems = 4;
constexpr long long kGNStreamBytesPerCta = 96 * 1024;
__global__ void __launch_bounds__(kGNThreads) gn_apply_stream_kernel(GNArgs a, int rows_per_cta) {
  pdl_wait();
  extern __shared__ float sm[];  // A[C], B[C], mean[G], rstd[G]
  float* s_A = sm;
  float* s_B = sm + a.C;
  float* s_mean = s_B + a.C;
  float* s_rstd = s_mean + a.G;
  const int b = blockIdx.y;
  const int vec_per_row = a.C >> 2;
  const long long p0 = (long long)blockIdx.x * rows_per_cta;
  const long long p1 = min(a.HW, p0 + (long long)rows_per_cta);
  const int n_items = (int)((p1 - p0) * vec_per_row);
  const long long row0 = (long long)b * a.HW + p0;
  // software pipeline: the loads of batch i+1 are in flight while batch i is normalised — and the loads of batch 0
  // while the CTA derives its affine table from the column statistics (a chain of dependent global loads that would
  // otherwise leave the CTA without a byte in flight for a third of its life: 3.2 -> see profiles/r02_hbm_kernels_*)
  float4 nxt[kGNStreamItems];
  auto issue = [&](int i0) {
#pragma unroll
    for (int k = 0; k < kGNStreamItems; ++k) {
      const int idx = i0 + k * kGNThreads;
      if (idx < n_items) {
        const int pr = idx / vec_per_row;
        const int cc = (idx - pr * vec_per_row) << 2;
        const long long row = row0 + pr;
        nxt[k] = cc < a.C1 ? __ldcs(reinterpret_cast<const float4*>(a.x1 + row * a.C1 + cc))
                           : __ldcs(reinterpret_cast<const float4*>(a.x2 + row * a.C2 + (cc - a.C1)));
      }
    }
  };
  issue(threadIdx.x);
  if (a.cs1) {
    gn_stats_from_colsums(a, b, s_mean, s_rstd);
    __syncthreads();
  }
  for (int c = threadIdx.x; c < a.C; c += kGNThreads) {
    const int g = c / a.cpg;
    const float2 mr = a.cs1 ? make_float2(s_mean[g], s_rstd[g])
                            : *reinterpret_cast<const float2*>(a.stats + ((long long)b * a.G + g) * 2);
    const float A = mr.y * __ldg(a.gamma + c);
    s_A[c] = A;
    s_B[c] = __ldg(a.beta + c) - mr.x * A;
  }
  __syncthreads();
  for (int i0 = threadIdx.x; i0 < n_items; i0 += kGNThreads * kGNStreamItems) {
    float4 v[kGNStreamItems];
#pragma unroll
    for (int k = 0; k < kGNStreamItems; ++k) v[k] = nxt[k];
    issue(i0 + kGNThreads * kGNStreamItems);
#pragma unroll
    for (int k = 0; k < kGNStreamItems; ++k) {
      const int idx = i0 + k * kGNThreads;
      if (idx >= n_items) continue;
      const int pr = idx / vec_per_row;
      const int c = (idx - pr * vec_per_row) << 2;
      const float in[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
      const float4 A4 = *reinterpret_cast<const float4*>(s_A + c);
      const float4 B4 = *reinterpret_cast<const float4*>(s_B + c);
      float o[4] = {fmaf(in[0], A4.x, B4.x), fmaf(in[1], A4.y, B4.y), fmaf(in[2], A4.z, B4.z), fmaf(in[3], A4.w, B4.w)};
      if (a.silu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = silu_f(o[e]);
      }
      const long long off = (row0 + pr) * a.C + c;
      op2_t h0 = ff2op2(o[0], o[1]);
      op2_t h1 = ff2op2(o[2], o[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(a.out + off) = pk;
      if (a.raw_out) {
        op2_t r0 = ff2op2(in[0], in[1]);
        op2_t r1 = ff2op2(in[2], in[3]);
        uint2 rk;
        rk.x = *reinterpret_cast<uint32_t*>(&r0);
        rk.y = *reinterpret_cast<uint32_t*>(&r1);
        *reinterpret_cast<uint2*>(a.raw_out + off) = rk;
      }
      if (a.cat_out) *reinterpret_cast<float4*>(a.cat_out + off) = v[k];
    }
  }
  pdl_trigger();
}

// one warp per row; the row is read once into registers: NV float4 per lane (C <= 128*NV), NV a template constant
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long rows, int C, float eps,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        op_t* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  // the affine parameters are weights (never written by the previous kernel): fetch them before the dependency wait
  // and before the row, so their latency is not a second round trip after the reductions
  float4 gm[NV], bt[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < C) {
      gm[k] = __ldg(reinterpret_cast<const float4*>(gamma + c));
      bt[k] = __ldg(reinterpret_cast<const float4*>(beta + c));
    }
  }
  pdl_wait();
  if (row >= rows) return;
  const float* xr = x + row * C;
  float4 v[NV];
  float su = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    v[k] = c < C ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    su += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) su += __shfl_xor_sync(0xffffffffu, su, o);
  pdl_trigger();   // the row is in registers: the rest of this kernel may overlap the next launch
  const float mean = su / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < C) {
      const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
      sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  op_t* orow = out + row * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < C) {
      const float4 g = gm[k];
      const float4 bb = bt[k];
      op2_t h0 = ff2op2((v[k].x - mean) * rstd * g.x + bb.x, (v[k].y - mean) * rstd * g.y + bb.y);
      op2_t h1 = ff2op2((v[k].z - mean) * rstd * g.z + bb.z, (v[k].w - mean) * rstd * g.w + bb.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(orow + c) = pk;
    }
  }
}

template <int NV>
cudaError_t launch_ln(const float* x, long long rows, int C, float eps, const float* gamma, const float* beta,
                      op_t* out, cudaStream_t st) {
  // one warp per row; few rows (reverse process at batch 2: 128-2048 rows) -> fewer rows per CTA so the rows spread
  // over all SMs instead of queueing on a few
  const int warps = rows >= 148 * 16 ? 4 : (rows >= 148 * 4 ? 2 : 1);
  return launch_kernel_family(4, layernorm_kernel<NV>, dim3((unsigned)ceil_div64(rows, warps)), dim3(warps * 32), (size_t)0, st, x,
                       rows, C, eps, gamma, beta, out);
}

}  // namespace
}  // namespace aedit

using namespace aedit;

static thread_local long long g_gn_stream_min_bytes = 8ll << 20;
extern "C" void ae_set_gn_stream_min_bytes(int64_t bytes) { g_gn_stream_min_bytes = bytes; }

extern "C" int64_t ae_groupnorm_workspace_bytes(int B, int groups) {
  // partial sums [B, kMaxSplits, G, 2] double + stats [B, G, 2] float + counters [B] u32 (must start zeroed)
  return (int64_t)B * kMaxSplits * groups * 2 * 8 + (int64_t)B * groups * 2 * 4 + (int64_t)B * 4 + 64;
}

static int groupnorm_impl(const float* x1, int C1, const float* x2, int C2, int B, int64_t HW, int groups, float eps,
                          const float* gamma, const float* beta, int silu, void* out_bf16, void* raw_out_bf16,
                          float* cat_out_f32, float* workspace, const int64_t* cs1, const int64_t* cs2, ae_stream stream) {
  AE_CHECK_ARG(x1 && C1 > 0 && B > 0 && HW > 0 && groups > 0, "ae_groupnorm: bad argument");
  AE_CHECK_ARG((x2 != nullptr) == (C2 > 0), "ae_groupnorm: x2/C2 mismatch");
  const int C = C1 + C2;
  AE_CHECK_ARG(C % groups == 0, "ae_groupnorm: C=%d not divisible by groups=%d", C, groups);
  AE_CHECK_ARG(C1 % 4 == 0 && C2 % 4 == 0, "ae_groupnorm: channel counts must be multiples of 4 (C1=%d C2=%d)", C1, C2);
  AE_CHECK_ARG(gamma && beta && out_bf16 && workspace, "ae_groupnorm: null pointer");
  AE_CHECK_ARG(HW * (int64_t)(C / 4) < 2147483647LL, "ae_groupnorm: sample too large (HW*C/4 >= 2^31)");
  AE_CHECK_ARG(C <= 2560, "ae_groupnorm: C=%d > 2560 channels", C);
  GNArgs a;
  a.x1 = x1;
  a.x2 = x2;
  a.C1 = C1;
  a.C2 = C2;
  a.C = C;
  a.G = groups;
  a.cpg = C / groups;
  a.B = B;
  a.HW = HW;
  a.cs1 = reinterpret_cast<const long long*>(cs1);
  a.cs2 = reinterpret_cast<const long long*>(cs2);
  // thread block: TX channel quads x TY positions (<= 512 threads, <= 4 quads per thread)
  const int nq = C / 4;
  int TX = nq < 64 ? nq : 64;
  while ((nq + TX - 1) / TX > kGNMaxQuadsPerThread) TX *= 2;
  AE_CHECK_ARG(TX <= 512, "ae_groupnorm: C=%d too wide", C);
  int TY = 512 / TX;
  if (TY > 16) TY = 16;
  if (TY < 1) TY = 1;
  a.eps = eps;
  a.gamma = gamma;
  a.beta = beta;
  a.silu = silu;
  a.out = reinterpret_cast<op_t*>(out_bf16);
  a.raw_out = reinterpret_cast<op_t*>(raw_out_bf16);
  a.cat_out = cat_out_f32;
  char* wsb = reinterpret_cast<char*>(workspace);
  a.partial = reinterpret_cast<double*>(wsb);
  a.stats = reinterpret_cast<float*>(wsb + (size_t)B * kMaxSplits * groups * 2 * 8);
  a.counters = reinterpret_cast<unsigned int*>(wsb + (size_t)B * kMaxSplits * groups * 2 * 8 + (size_t)B * groups * 2 * 4);
  // position splits: enough CTAs to cover the machine at small batch, never more than kMaxSplits per sample
  int S = (int)ceil_div64(HW, 4 * TY);
  const int want = (296 + B - 1) / B;
  if (S > want) S = want;
  if (S > kMaxSplits) S = kMaxSplits;
  if (S < 1) S = 1;
  a.chunk = ceil_div64(HW, S);
  a.S = (int)ceil_div64(HW, a.chunk);
  cudaStream_t st = as_stream(stream);
  if (!cs1) {
    const size_t smem = (size_t)TY * 2 * C * sizeof(float);
    static size_t smem_set = 48 * 1024;
    if (smem > smem_set) {
      cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      smem_set = 160 * 1024;
    }
    AE_CHECK_ARG(smem <= 160 * 1024, "ae_groupnorm: C=%d needs too much shared memory", C);
    if (!(g_skip_mask & 4)) launch_kernel_family(1, gn_stats_kernel, dim3(a.S, B), dim3(TX, TY), smem, st, a);
  }
  if (!cs1) {
    int rc = launched("ae_groupnorm(stats)");
    if (rc) return rc;
  }
  {
    const long long items = HW * (C / 4);
    const unsigned gx = (unsigned)ceil_div64(items, (long long)kGNThreads * kGNItems);
    // large tensors (forward-process chunks): streaming variant — the one-round kernel keeps ~100 registers per
    // thread (2 CTAs per SM) and reaches ~2 TB/s; small ones (reverse process) are latency-bound and keep it
    const long long bytes = (long long)B * HW * C * 4;
    if (g_skip_mask & 8) {
    } else if (bytes >= g_gn_stream_min_bytes) {
      const long long row_bytes = (long long)C * 4;
      long long rows_per_cta = (kGNStreamBytesPerCta + row_bytes - 1) / row_bytes;
      if (rows_per_cta > HW) rows_per_cta = HW;
      const unsigned sx = (unsigned)ceil_div64(HW, rows_per_cta);
      launch_kernel_family(2, gn_apply_stream_kernel, dim3(sx, B), dim3(kGNThreads),
                           (size_t)((2 * C + 2 * groups) * sizeof(float)), st, a, (int)rows_per_cta);
    } else {
      launch_kernel_family(2, gn_apply_kernel, dim3(gx, B), dim3(kGNThreads),
                           (size_t)((2 * groups + 2 * C) * sizeof(float)), st, a);
    }
  }
  return launched("ae_groupnorm(apply)");
}

extern "C" int ae_groupnorm(const float* x1, int C1, const float* x2, int C2, int B, int64_t HW, int groups, float eps,
                            const float* gamma, const float* beta, int silu, void* out_bf16, void* raw_out_bf16,
                            float* cat_out_f32, float* workspace, ae_stream stream) {
  return groupnorm_impl(x1, C1, x2, C2, B, HW, groups, eps, gamma, beta, silu, out_bf16, raw_out_bf16, cat_out_f32,
                        workspace, nullptr, nullptr, stream);
}

extern "C" int ae_groupnorm_cs(const float* x1, int C1, const int64_t* colstats1, const float* x2, int C2,
                               const int64_t* colstats2, int B, int64_t HW, int groups, float eps, const float* gamma,
                               const float* beta, int silu, void* out_bf16, void* raw_out_bf16, float* cat_out_f32,
                               float* workspace, ae_stream stream) {
  AE_CHECK_ARG(colstats1 && ((colstats2 != nullptr) == (C2 > 0)), "ae_groupnorm_cs: column statistics missing");
  AE_CHECK_ARG((reinterpret_cast<uintptr_t>(colstats1) & 15) == 0 && (reinterpret_cast<uintptr_t>(colstats2) & 15) == 0,
               "ae_groupnorm_cs: column statistics must be 16-byte aligned");
  return groupnorm_impl(x1, C1, x2, C2, B, HW, groups, eps, gamma, beta, silu, out_bf16, raw_out_bf16, cat_out_f32,
                        workspace, colstats1, colstats2, stream);
}

extern "C" int ae_layernorm(const float* x, int64_t rows, int C, float eps, const float* gamma, const float* beta,
                            void* out_bf16, ae_stream stream) {
  AE_CHECK_ARG(x && gamma && beta && out_bf16 && rows > 0 && C > 0, "ae_layernorm: bad argument");
  AE_CHECK_ARG(C % 4 == 0 && C <= 2048, "ae_layernorm: C=%d must be a multiple of 4 and <= 2048", C);
  op_t* o = reinterpret_cast<op_t*>(out_bf16);
  cudaStream_t st = as_stream(stream);
  const int nv = (C + 127) / 128;
  cudaError_t e;
  if (g_skip_mask & 16) return AE_OK;
  if (nv <= 1) e = launch_ln<1>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 2) e = launch_ln<2>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 3) e = launch_ln<3>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 4) e = launch_ln<4>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 5) e = launch_ln<5>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 6) e = launch_ln<6>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 8) e = launch_ln<8>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 10) e = launch_ln<10>(x, rows, C, eps, gamma, beta, o, st);
  else if (nv <= 12) e = launch_ln<12>(x, rows, C, eps, gamma, beta, o, st);
  else e = launch_ln<16>(x, rows, C, eps, gamma, beta, o, st);
  if (e != cudaSuccess) return fail(AE_ECUDA, "ae_layernorm launch: %s", cudaGetErrorString(e));
  return launched("ae_layernorm");
}
