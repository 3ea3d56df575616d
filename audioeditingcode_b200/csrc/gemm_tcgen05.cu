// tcgen05 / TMA GEMM and implicit-GEMM convolution (K1/K2/K5-linear of SURVEY.md §2.2).
//
//   D[M,N] = alpha * A[M,K] · W[N,K]^T (+bias)(+rowbias)(+residual) -> act -> {f32, bf16} outputs
//
// Replaces every cuDNN/cuBLAS call the reference's U-Net evaluation lands in (conv3x3 / conv1x1 / Linear of
// diffusers ResnetBlock2D + Transformer2DModel driven from code/models.py:293-388; in-tree twins
// code/audioldm/latent_diffusion/openaimodel.py:213-244, attention.py:220-323).
//
// Tile: 128 (M) x BN (N) x 64 (K) of 16-bit operands (op_t: fp16; bf16 with -DAE_OPERAND_BF16), fp32 accumulators in TMEM.
//   gemm_tcgen05_kernel: one CTA per output tile, 6 warps (gemm_persistent_kernel further down: one CTA per SM)
//   warp 0  TMA producer (cp.async.bulk.tensor, 128B-swizzled K-major tiles, 2..6-stage mbarrier ring; short rings keep 2-3 CTAs per SM)
//   warp 1  TMEM allocator + single-thread tcgen05.mma issuer (4 UMMAs of K=16 per stage) + tcgen05.commit
//   warps 2-5  epilogue: tcgen05.ld; fused bias / time-embedding row bias / residual / SiLU / GEGLU / GroupNorm column
//              statistics; row-per-thread stores, or (fp32 outputs) a shared-memory transpose for row-contiguous accesses
// Implicit convolution: the A operand is a channels-last image [B,H,W,C]; K-block kb = (tap, 64-channel slab);
// the producer issues a 4-D TMA box {64 ch, Wb, Hb, Bb} at (c0, w0+dw, h0+dh, b0) — out-of-bounds rows/cols are
// zero-filled by TMA, which is exactly the convolution's zero padding.  The 128 tile rows are the flattened
// (b,h,w) positions m0..m0+127, so the epilogue is identical to the plain GEMM.
// Small-M layers (deep U-Net levels at batch 2: one or two M tiles, K up to 8640): split-K over blockIdx.z, fp32
// partial tiles to a workspace, then a fully parallel fixed-order reduce + epilogue kernel (no atomics).
// Determinism / batch invariance: the split factor depends only on (tiles, K), never on data; a sample's reduction
// order never depends on what else is in the batch except through the tile count (documented in DESIGN.md §4).
#include <cuda.h>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace aedit {
// Launch settings are PER HOST THREAD (include/aedit.h, "Settings"): the host toggles them around CUDA-graph captures,
// and two threads driving different streams must not see each other's toggles.
thread_local int g_use_pdl = 2;
thread_local int g_launch_priority = 0;
thread_local int g_skip_mask = 0;
thread_local int g_pdl_extra = 0;
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kATileBytes = BM * BK * 2;  // 16 KiB

struct GemmDev {
  int M, N;
  int num_kblocks;
  int kb_per_split;  // == num_kblocks when not split
  int split;         // 1: blockIdx.z is a K split (partials to out_f32 + z*stride_out), else a batch index
  // implicit conv
  int conv, H, W, HW, cblocks, kw, dil_h, dil_w, pad_h, pad_w;
  // epilogue
  const float* bias;
  const float* rowbias;
  long long ld_rowbias;
  int rows_per_group;
  const float* residual;
  long long ld_res;
  float* out_f32;
  long long ld_out_f32;
  op_t* out_bf16;
  long long ld_out_bf16;
  long long stride_out, stride_res;
  int w_dynamic;  // 1: the W operand is produced by a preceding kernel (never prefetch it ahead of the dependency)
  int m_in_x;    // 1: M tiles on blockIdx.x (only when there are more than 65535 of them), else N tiles (default)
  int fast_epi;  // 1: operands / outputs are 16-byte tileable -> coalesced staged epilogue (epilogue_strip)
  int act;  // 0 none, 1 SiLU, 2 GEGLU (output width N/2: out[16q+i] = acc[32q+i] * gelu(acc[32q+16+i]))
  float alpha;
  // GroupNorm column statistics of the OUTPUT (see ae_gemm_args.colstats): fixed-point accumulators [sample][N][2]
  unsigned long long* colstats;
  int cs_rows;   // rows per sample (multiple of 32)
  // act == 3: grouped softmax over the output columns (cross-attention against frozen text folded into two GEMMs,
  // see ae_gemm_args.sm_*): column n = (text_row * heads + head) * sm_L + key
  int sm_L, sm_block, sm_rows;
  const int* sm_slot;
  const float* sm_bias;
};

// softmax over groups of L consecutive accumulator columns of one output row (one (text row, head) each); groups of
// other text rows than the sample's own are zeroed.  `bias` (optional): additive, per (text row, key).
template <int L>
__device__ __forceinline__ void softmax_groups(float (&acc)[32], int nbase, int block, int my_row, const float* bias) {
#pragma unroll
  for (int g = 0; g < 32 / L; ++g) {
    const int r = (nbase + g * L) / block;
    if (r != my_row) {
#pragma unroll
      for (int l = 0; l < L; ++l) acc[g * L + l] = 0.f;
      continue;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      float v = acc[g * L + l];
      if (bias) v += __ldg(bias + r * L + l);
      acc[g * L + l] = v;
      mx = fmaxf(mx, v);
    }
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((acc[g * L + l] - mx) * 1.4426950408889634f));
      acc[g * L + l] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int l = 0; l < L; ++l) acc[g * L + l] *= inv;
  }
}

// Fixed-point scales of the column statistics: sum * 2^28, sum of squares * 2^24, accumulated with integer atomics —
// integer addition is associative, so the totals do not depend on the order in which CTAs arrive (deterministic),
// unlike floating-point atomics.  Headroom: |sum| < 3.4e10, sum of squares < 5.5e11 per (sample, channel).
constexpr double kCsScaleSum = 268435456.0;
constexpr double kCsScaleSq = 16777216.0;
__device__ __forceinline__ void colstats_add(unsigned long long* cs, long long sample, int N, int col, float s, float q) {
  unsigned long long* dst = cs + ((sample * N + col) << 1);
  atomicAdd(dst, (unsigned long long)__double2ll_rn((double)s * kCsScaleSum));
  atomicAdd(dst + 1, (unsigned long long)__double2ll_rn((double)q * kCsScaleSq));
}

// STAGES = 3: 2-3 CTAs per SM (large grids: one CTA's epilogue overlaps another's main loop).
// STAGES = 6: grids that cannot fill the machine anyway (one CTA per SM) need the deeper ring to cover TMA latency.
template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kBTileBytes = BN * BK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kBarBytes = 128;
  static constexpr int kEpiBytes = 2 * BN * 4;  // bias / row-bias values of the tile (row-per-thread epilogue)
  // the staged epilogue transposes the four 32-row strips through the (then idle) ring: pitch BN+4 floats
  static constexpr int kStripBytes = 4 * 32 * (BN + 4) * 4;
  static constexpr int kRingBytes = STAGES * kStageBytes > kStripBytes ? STAGES * kStageBytes : ((kStripBytes + 1023) / 1024) * 1024;
  static constexpr int kCsBytes = BN * 16 + 16;   // column-statistics accumulators of the tile (u64 sum, sumsq) + ticket
  static constexpr int kTotal = kRingBytes + kBarBytes + kEpiBytes + kCsBytes + 1024;  // +1024 manual alignment slack
};

// gelu(x) = x * Phi(x) (erf form, attention.py:37-44 / torch F.gelu default).  Branch-free: erf(|z|), z = x / sqrt(2),
// from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z);
// for x < 0, Phi = 0.5 * poly * exp(-z^2) directly (no cancellation in the tail).  libdevice's erff has two regimes, so
// a warp whose lanes straddle |z| ~ 0.9 executes both; the GEGLU epilogue is ALU-bound (one value per 2 accumulators).
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float pl = fmaf(1.061405429f, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  const float half_tail = 0.5f * pl * t * e;            // 0.5 * erfc(|z|)
  return x * (x >= 0.f ? 1.0f - half_tail : half_tail);
}

// Fused epilogue of one 32-column chunk of one output row: bias, time-embedding row bias, residual, activation,
// fp32 / bf16 stores.  `acc` already holds alpha * accumulator.
// `sb` / `srb` (optional): the tile's bias / row-bias values staged in shared memory by the caller (indexed from the
// chunk's first column) — fetched once per CTA before the accumulator is ready instead of once per chunk after it.
__device__ __forceinline__ void epilogue_chunk(const GemmDev& p, float (&acc)[32], long long m, int nbase, int z,
                                               const float* sb = nullptr, const float* srb = nullptr) {
  if (m >= p.M || nbase >= p.N) return;
  const int nvalid = min(32, p.N - nbase);
  if (p.bias && sb) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += sb[j];     // columns >= N hold 0
  } else if (p.bias) {
    if (nvalid == 32) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + nbase);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(b4 + j);
        acc[4 * j + 0] += t.x;
        acc[4 * j + 1] += t.y;
        acc[4 * j + 2] += t.z;
        acc[4 * j + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) acc[j] += __ldg(p.bias + nbase + j);
    }
  }
  if (p.rowbias && srb) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += srb[j];
  } else if (p.rowbias) {
    const float* rb = p.rowbias + (m / p.rows_per_group) * p.ld_rowbias;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) acc[j] += __ldg(rb + nbase + j);
  }
  if (p.residual) {
    const float* res = p.residual + (long long)z * p.stride_res + m * p.ld_res;
    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(res + nbase) & 15) == 0)) {
      const float4* r4 = reinterpret_cast<const float4*>(res + nbase);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = r4[j];
        acc[4 * j + 0] += t.x;
        acc[4 * j + 1] += t.y;
        acc[4 * j + 2] += t.z;
        acc[4 * j + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) acc[j] += res[nbase + j];
    }
  }
  int obase = nbase, ovalid = nvalid;
  if (p.act == 3) {
    const int my_row = __ldg(p.sm_slot + m / p.sm_rows);
    if (p.sm_L == 8) softmax_groups<8>(acc, nbase, p.sm_block, my_row, p.sm_bias);
    else if (p.sm_L == 16) softmax_groups<16>(acc, nbase, p.sm_block, my_row, p.sm_bias);
    else softmax_groups<32>(acc, nbase, p.sm_block, my_row, p.sm_bias);
  } else if (p.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = silu_f(acc[j]);
  } else if (p.act == 2) {
    // weight rows were interleaved host-side: columns [32q, 32q+16) = value, [32q+16, 32q+32) = gate
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = acc[j] * gelu_erf_f(acc[16 + j]);
    obase = nbase >> 1;
    ovalid = nvalid >> 1;
  }
  if (p.out_f32) {
    float* of = p.out_f32 + (long long)z * p.stride_out + m * p.ld_out_f32;
    if (ovalid == 32 && ((reinterpret_cast<uintptr_t>(of + obase) & 15) == 0)) {
      float4* o4 = reinterpret_cast<float4*>(of + obase);
#pragma unroll
      for (int j = 0; j < 8; ++j) o4[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ovalid) of[obase + j] = acc[j];
    }
  }
  if (p.out_bf16) {
    op_t* ob = p.out_bf16 + (long long)z * p.stride_out + m * p.ld_out_bf16;
    if ((ovalid == 32 || ovalid == 16) && ((reinterpret_cast<uintptr_t>(ob + obase) & 15) == 0)) {
      uint4* o4 = reinterpret_cast<uint4*>(ob + obase);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (8 * j < ovalid) {
          op2_t h0 = ff2op2(acc[8 * j + 0], acc[8 * j + 1]);
          op2_t h1 = ff2op2(acc[8 * j + 2], acc[8 * j + 3]);
          op2_t h2 = ff2op2(acc[8 * j + 4], acc[8 * j + 5]);
          op2_t h3 = ff2op2(acc[8 * j + 6], acc[8 * j + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          o4[j] = pk;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ovalid) ob[obase + j] = f2op(acc[j]);
    }
  }
}

// ---- coalesced epilogue of one warp's 32-row strip of the output tile -------------------------------------------
// tcgen05.ld hands every thread one accumulator ROW (TMEM lane); writing that row straight to global memory makes
// each warp instruction touch 32 different rows, 16 bytes each (the pre-r9 epilogue: 2.1 TB/s on the residual
// linears, one exposed L2 round trip per 32-column chunk).  Here the strip is transposed through shared memory (the
// TMA ring is idle once the accumulator is complete; row pitch BN+4 floats keeps both the row-owner float4 writes and
// the row-contiguous float4 reads conflict-free), and every global access — residual read, time-embedding row bias,
// fp32 / bf16 stores — is a row-contiguous 16 bytes per lane (BN/4 lanes per row, 128/BN rows per instruction).
// The bias / row-bias loads, an L2 prefetch of the strip's residual and the residual loads of the first PF iterations
// are issued BEFORE the accumulator barrier is waited on, so their latency overlaps the main loop; later batches are
// loaded one batch ahead.  The iteration loops are deliberately NOT fully unrolled: the epilogue runs once per CTA, so
// its instructions are fetched cold — a 32x unrolled body (v10a) cost +10 us per launch in instruction-cache misses.
template <int BN>
__device__ __forceinline__ void epilogue_strip(const GemmDev& p, float* strip, uint32_t tmem_strip, uint64_t* tmem_full_bar,
                                               long long m_base, int n0, int zo, int lane, uint32_t parity = 0,
                                               uint64_t* tmem_release_bar = nullptr,
                                               unsigned long long* cs_tile = nullptr) {
  constexpr int PITCH = BN + 4;
  constexpr int LPR = BN / 4;      // lanes per output row
  constexpr int RPI = 32 / LPR;    // rows per iteration
  constexpr int NIT = 32 / RPI;    // iterations per strip
  constexpr int PF = BN == 128 ? 8 : 4;   // iterations per residual batch (one batch of loads in flight)
  constexpr int NB = NIT / PF;
  const int sub = lane / LPR, c4 = lane % LPR;
  const bool geglu = p.act == 2;
  // column(s) owned by this lane: plain -> 4 consecutive columns; GEGLU -> 2 value + 2 gate columns of a 32-block
  const int oc = c4 * 2, q = oc >> 4, off = oc & 15;
  const int ncol = geglu ? (32 * q + off) : c4 * 4;       // tile-local column of the first (value) element
  const int n = n0 + ncol;
  const bool col_ok = geglu ? (n0 + 32 * q + 16 + off + 1 < p.N) : (n + 3 < p.N);
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);            // plain: bias[n..n+3]; GEGLU: value bias x,y  gate bias z,w
  if (p.bias && col_ok) {
    if (geglu) {
      const float2 bv = *reinterpret_cast<const float2*>(p.bias + n);
      const float2 bg = *reinterpret_cast<const float2*>(p.bias + n + 16);
      b4 = make_float4(bv.x, bv.y, bg.x, bg.y);
    } else {
      b4 = *reinterpret_cast<const float4*>(p.bias + n);
    }
  }
  // time-embedding row bias: one value per (row group, column); constant over the strip unless a group ends inside
  const bool rb_const = p.rowbias && (m_base / p.rows_per_group) == ((m_base + 31) / p.rows_per_group);
  float4 rb4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rb_const && col_ok && m_base < p.M)
    rb4 = *reinterpret_cast<const float4*>(p.rowbias + (m_base / p.rows_per_group) * p.ld_rowbias + n);
  const float* res_base = p.residual ? p.residual + (long long)zo * p.stride_res + n : nullptr;
  if (res_base && m_base + lane < p.M) {
    // pull the whole strip of the residual towards L2 (row `lane`, BN*4 bytes) while the main loop runs ...
    const char* rp = reinterpret_cast<const char*>(p.residual + (long long)zo * p.stride_res + (m_base + lane) * p.ld_res + n0);
#pragma unroll
    for (int l = 0; l < BN * 4 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + l * 128));
  }
  float4 cur[PF];   // ... and the first batch into registers
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const long long m = m_base + (long long)i * RPI + sub;
    cur[i] = (res_base && col_ok && m < p.M) ? *reinterpret_cast<const float4*>(res_base + m * p.ld_res)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  ptx::mbar_wait(tmem_full_bar, parity);
  ptx::tcgen05_fence_after();
  __syncwarp();   // persistent kernel: every lane is done reading the previous tile's strip
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(tmem_strip + c * 32, v);
    ptx::tmem_ld_wait();
    float4* dst = reinterpret_cast<float4*>(strip + lane * PITCH + c * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      dst[j] = make_float4(__uint_as_float(v[4 * j]) * p.alpha, __uint_as_float(v[4 * j + 1]) * p.alpha,
                           __uint_as_float(v[4 * j + 2]) * p.alpha, __uint_as_float(v[4 * j + 3]) * p.alpha);
  }
  __syncwarp();
  if (tmem_release_bar) {
    // persistent kernel: the accumulator buffer is in shared memory now, hand it back to the MMA warp
    ptx::tcgen05_fence_before();
    if (lane == 0) ptx::mbar_arrive(tmem_release_bar);
  }
  const unsigned cs_mask = __ballot_sync(0xffffffffu, col_ok);   // lanes that stay (whole column classes)
  if (!col_ok) return;
  if (geglu) {
    const long long ocol = (n0 >> 1) + oc;
#pragma unroll 1
    for (int it = 0; it < NIT; ++it) {
      const int rl = it * RPI + sub;
      const long long m = m_base + rl;
      if (m >= p.M) break;
      const float* srow = strip + rl * PITCH;
      const float2 va = *reinterpret_cast<const float2*>(srow + ncol);
      const float2 ga = *reinterpret_cast<const float2*>(srow + ncol + 16);
      const float o0 = (va.x + b4.x) * gelu_erf_f(ga.x + b4.z);
      const float o1 = (va.y + b4.y) * gelu_erf_f(ga.y + b4.w);
      if (p.out_f32)
        *reinterpret_cast<float2*>(p.out_f32 + (long long)zo * p.stride_out + m * p.ld_out_f32 + ocol) = make_float2(o0, o1);
      if (p.out_bf16) {
        op2_t h = ff2op2(o0, o1);
        *reinterpret_cast<op2_t*>(p.out_bf16 + (long long)zo * p.stride_out + m * p.ld_out_bf16 + ocol) = h;
      }
    }
    return;
  }
  float* of = p.out_f32 ? p.out_f32 + (long long)zo * p.stride_out + n : nullptr;
  op_t* ob = p.out_bf16 ? p.out_bf16 + (long long)zo * p.stride_out + n : nullptr;
  // column statistics of this strip (rows in increasing order per lane, then a fixed tree over the lanes that share
  // the columns): per-column sum and sum of squares of the final fp32 values
  float4 cs_s = make_float4(0.f, 0.f, 0.f, 0.f), cs_q = cs_s;
#pragma unroll 1
  for (int b = 0; b < NB; ++b) {
    float4 r[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) r[i] = cur[i];
    if (b + 1 < NB) {
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const long long m = m_base + (long long)((b + 1) * PF + i) * RPI + sub;
        cur[i] = (res_base && m < p.M) ? *reinterpret_cast<const float4*>(res_base + m * p.ld_res)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int rl = (b * PF + i) * RPI + sub;
      const long long m = m_base + rl;
      if (m < p.M) {
        float4 a = *reinterpret_cast<const float4*>(strip + rl * PITCH + ncol);
        a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
        if (p.rowbias) {
          float4 t = rb4;
          if (!rb_const) t = *reinterpret_cast<const float4*>(p.rowbias + (m / p.rows_per_group) * p.ld_rowbias + n);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        a.x += r[i].x; a.y += r[i].y; a.z += r[i].z; a.w += r[i].w;
        if (p.act == 1) {
          a.x = silu_f(a.x); a.y = silu_f(a.y); a.z = silu_f(a.z); a.w = silu_f(a.w);
        }
        if (p.colstats) {
          cs_s.x += a.x; cs_s.y += a.y; cs_s.z += a.z; cs_s.w += a.w;
          cs_q.x = fmaf(a.x, a.x, cs_q.x); cs_q.y = fmaf(a.y, a.y, cs_q.y);
          cs_q.z = fmaf(a.z, a.z, cs_q.z); cs_q.w = fmaf(a.w, a.w, cs_q.w);
        }
        if (of) *reinterpret_cast<float4*>(of + m * p.ld_out_f32) = a;
        if (ob) {
          op2_t h0 = ff2op2(a.x, a.y);
          op2_t h1 = ff2op2(a.z, a.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(ob + m * p.ld_out_bf16) = pk;
        }
      }
    }
  }
  if (p.colstats) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      cs_s.x += __shfl_xor_sync(cs_mask, cs_s.x, o); cs_s.y += __shfl_xor_sync(cs_mask, cs_s.y, o);
      cs_s.z += __shfl_xor_sync(cs_mask, cs_s.z, o); cs_s.w += __shfl_xor_sync(cs_mask, cs_s.w, o);
      cs_q.x += __shfl_xor_sync(cs_mask, cs_q.x, o); cs_q.y += __shfl_xor_sync(cs_mask, cs_q.y, o);
      cs_q.z += __shfl_xor_sync(cs_mask, cs_q.z, o); cs_q.w += __shfl_xor_sync(cs_mask, cs_q.w, o);
    }
    const long long sample = m_base / p.cs_rows;       // a 32-row strip never straddles samples (cs_rows % 32 == 0)
    if (cs_tile == nullptr) {
      if (sub == 0 && m_base < p.M) {
        colstats_add(p.colstats, sample, p.N, n + 0, cs_s.x, cs_q.x);
        colstats_add(p.colstats, sample, p.N, n + 1, cs_s.y, cs_q.y);
        colstats_add(p.colstats, sample, p.N, n + 2, cs_s.z, cs_q.z);
        colstats_add(p.colstats, sample, p.N, n + 3, cs_s.w, cs_q.w);
      }
    } else {
      // the tile's four strips belong to one sample: combine them in shared memory first (fixed-point integer adds are
      // exact, so the order of the warps does not matter) and let the last strip to arrive issue ONE global atomic per
      // column and moment — 4x fewer L2 atomics than per strip (they cost 5-10 % of a B = 100 GEMM, ncu v31)
      if (sub == 0) {
        const float sv[4] = {cs_s.x, cs_s.y, cs_s.z, cs_s.w}, qv[4] = {cs_q.x, cs_q.y, cs_q.z, cs_q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          atomicAdd(cs_tile + 2 * (ncol + e), (unsigned long long)__double2ll_rn((double)sv[e] * kCsScaleSum));
          atomicAdd(cs_tile + 2 * (ncol + e) + 1, (unsigned long long)__double2ll_rn((double)qv[e] * kCsScaleSq));
        }
      }
      __threadfence_block();
      __syncwarp(cs_mask);
      unsigned ticket = 0;
      if (lane == 0) ticket = (unsigned)atomicAdd(cs_tile + 2 * BN, 1ull);
      ticket = __shfl_sync(cs_mask, ticket, 0);
      if (ticket == 3) {
        __threadfence_block();
        if (sub == 0) {       // the lanes that own the tile's valid columns (the others left at the column check)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            unsigned long long* dst = p.colstats + ((sample * p.N + n + e) << 1);
            atomicAdd(dst, cs_tile[2 * (ncol + e)]);
            atomicAdd(dst + 1, cs_tile[2 * (ncol + e) + 1]);
          }
        }
      }
    }
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  using L = SmemLayout<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms (TMA and UMMA both derive the XOR from address bits)
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kATileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kRingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + L::kRingBytes + L::kBarBytes);
  float* s_rowb = s_bias + BN;
  unsigned long long* s_cs = reinterpret_cast<unsigned long long*>(s_rowb + BN);   // [2*BN] sums + [1] ticket

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // N tiles vary fastest in launch order: the CTAs that share one 128-row A tile are co-resident and the tile is
  // fetched from HBM once (W is small and stays in L2 either way).  ncu on the r9 build (M fastest) showed
  // dram__bytes_read = 2.6x the algorithmic bytes for [102400 x 1536] x [384 x 1536]: A re-read once per N tile.
  const int tile_m = p.m_in_x ? blockIdx.x : blockIdx.y;
  const int tile_n = p.m_in_x ? blockIdx.y : blockIdx.x;
  const int z = blockIdx.z;
  const int m0 = tile_m * BM;
  const int n0 = tile_n * BN;
  const bool ksplit = p.split != 0;
  const int zb = ksplit ? 0 : z;  // batch coordinate of the tensor maps
  const int kb0 = ksplit ? z * p.kb_per_split : 0;
  const int kb1 = ksplit ? min(p.num_kblocks, kb0 + p.kb_per_split) : p.num_kblocks;
  const int zo = z;  // output / residual batch-or-partial index

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  if (p.colstats && warp >= 2)
    for (int i = (int)threadIdx.x - 64; i < 2 * BN + 1; i += kThreads - 64) s_cs[i] = 0ull;
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above overlapped the previous kernel (programmatic dependent launch).  The weight operand W never
  // depends on the previous kernel, so the producer also streams the first ring of W tiles before blocking; all
  // other global traffic (activations A, residual, outputs) waits for the dependency.
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int b0 = 0, h0 = 0, w0 = 0;
      if (p.conv) {
        b0 = m0 / p.HW;
        const int rem = m0 - b0 * p.HW;
        h0 = rem / p.W;
        w0 = rem - h0 * p.W;
      }
      const int npre = p.w_dynamic ? 0 : min(STAGES, kb1 - kb0);
      for (int s = 0; s < npre; ++s) {
        ptx::mbar_expect_tx(&full_bar[s], L::kStageBytes);
        ptx::tma_load_3d(&tmB, &full_bar[s], sB + s * L::kBTileBytes, (kb0 + s) * BK, n0, zb);
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const bool pre = (kb - kb0) < npre;   // first ring: slot known free, barrier armed, W already in flight
        if (!pre) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
        }
        if (p.conv) {
          const int tap = kb / p.cblocks;
          const int cb = kb - tap * p.cblocks;
          const int i = tap / p.kw;
          const int j = tap - i * p.kw;
          ptx::tma_load_4d(&tmA, &full_bar[stage], sA + stage * kATileBytes, cb * BK, w0 + j * p.dil_w - p.pad_w,
                           h0 + i * p.dil_h - p.pad_h, b0);
        } else {
          ptx::tma_load_3d(&tmA, &full_bar[stage], sA + stage * kATileBytes, kb * BK, m0, zb);
        }
        if (!pre) ptx::tma_load_3d(&tmB, &full_bar[stage], sB + stage * L::kBTileBytes, kb * BK, n0, zb);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tcgen05_fence_after();
        const uint32_t a_addr = ptx::smem_u32(sA + stage * kATileBytes);
        const uint32_t b_addr = ptx::smem_u32(sB + stage * L::kBTileBytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128 B swizzle atom
          const uint64_t da = ptx::umma_desc_k_sw128(a_addr + k * 32);
          const uint64_t db = ptx::umma_desc_k_sw128(b_addr + k * 32);
          ptx::umma_bf16_ss(tmem_base, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        // frees the smem slot when these MMAs retire
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      ptx::umma_commit(tmem_full_bar);  // accumulator complete
      // every MMA of this CTA is issued: let the next kernel of the stream be scheduled now, so that its launch
      // latency and prologue overlap this CTA's epilogue (it still waits for this grid's completion before reading)
      pdl_trigger();
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const long long m = (long long)m0 + row;
    pdl_wait();
    if (p.fast_epi) {
      float* strip = reinterpret_cast<float*>(smem) + (size_t)quad * 32 * (BN + 4);
      // column statistics: combine the four strips in shared memory when the whole tile lies in one sample
      const bool cs_one = p.colstats && (m0 / p.cs_rows) == ((m0 + BM - 1) / p.cs_rows) && (long long)m0 + BM <= p.M;
      epilogue_strip<BN>(p, strip, tmem_base + (static_cast<uint32_t>(quad * 32) << 16), tmem_full_bar,
                                       (long long)m0 + quad * 32, n0, zo, lane, 0, nullptr, cs_one ? s_cs : nullptr);
      ptx::tcgen05_fence_before();
    } else {
    const bool row_ok = m < p.M;
    if (p.residual && row_ok) {
      // pull this row's residual segment towards L2 while the main loop runs (BN*4 bytes = up to 4 lines)
      const char* rp = reinterpret_cast<const char*>(p.residual + (long long)zo * p.stride_res + m * p.ld_res + n0);
#pragma unroll
      for (int l = 0; l < BN * 4 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + l * 128));
    }
    // bias / row bias of the tile -> shared memory, one coalesced fetch per CTA while the main loop runs (the row bias
    // only when the whole tile lies in one row group, which is the rule for the time-embedding bias of the ResBlocks)
    const int etid = (int)threadIdx.x - 64;
    const bool stage_b = p.bias != nullptr;
    const long long mlast = min((long long)m0 + BM - 1, (long long)p.M - 1);
    const bool stage_rb = p.rowbias != nullptr && (m0 / p.rows_per_group) == (mlast / p.rows_per_group);
    if (stage_b || stage_rb) {
      for (int i = etid; i < BN; i += 128) {
        const bool ok = n0 + i < p.N;
        if (stage_b) s_bias[i] = ok ? __ldg(p.bias + n0 + i) : 0.f;
        if (stage_rb) s_rowb[i] = ok ? __ldg(p.rowbias + (m0 / p.rows_per_group) * p.ld_rowbias + n0 + i) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
      ptx::tmem_ld_wait();
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]) * p.alpha;
      epilogue_chunk(p, acc, m, n0 + c * 32, zo, stage_b ? s_bias + c * 32 : nullptr, stage_rb ? s_rowb + c * 32 : nullptr);
    }
    ptx::tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant for multi-wave grids (forward-process chunks): one CTA per SM walks the output tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ... (N tiles fastest, so the CTAs that share an A tile run at the same
// time).  The accumulator is double-buffered in TMEM (2 x BN columns): the MMA warp starts tile i+1 while the epilogue
// warps drain tile i, the TMA ring keeps streaming across tile boundaries, and the per-CTA fixed costs (launch, TMEM
// allocation, barrier init, pipeline ramp) are paid once per SM instead of once per tile.  Same arithmetic per
// output element as gemm_tcgen05_kernel (same K order, same epilogue code) -> same bits.
template <int BN, int STAGES, bool STAGED>
struct PersistLayout {
  static constexpr int kBTileBytes = BN * BK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kRingBytes = STAGES * kStageBytes;
  static constexpr int kStripBytes = STAGED ? ((4 * 32 * (BN + 4) * 4 + 1023) / 1024) * 1024 : 0;
  static constexpr int kBarBytes = 256;
  static constexpr int kEpiBytes = 2 * 2 * BN * 4;   // bias / row-bias values, one set per accumulator buffer
  static constexpr int kTotal = kRingBytes + kStripBytes + kBarBytes + kEpiBytes + 1024;
};

// Epilogue warps of the persistent kernel: the row-per-thread epilogue (bf16 / GEGLU outputs) is ALU- and latency-bound
// — the GEGLU layers execute ~20 instructions per accumulator on ONE warp per SM sub-partition (ncu r01 v35: 9 % of the
// warp slots) — so it runs on EIGHT warps: warps 2..5 and 6..9 share the four TMEM lane quadrants and split the tile's
// columns in halves.  The staged epilogue (fp32 outputs) keeps four.
template <bool STAGED>
constexpr int persist_threads() { return STAGED ? kThreads : kThreads + 128; }

template <int BN, int STAGES, bool STAGED>
__global__ void __launch_bounds__(persist_threads<STAGED>(), 1)
gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p,
                       int tiles_n, int n_tiles) {
  using L = PersistLayout<BN, STAGES, STAGED>;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kATileBytes;
  float* strips = reinterpret_cast<float*>(smem + L::kRingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kRingBytes + L::kStripBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + L::kRingBytes + L::kStripBytes + L::kBarBytes);   // [2][BN]
  float* s_rowb = s_bias + 2 * BN;                                                                  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = p.num_kblocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full_bar[s], 1);
      ptx::mbar_init(&tmem_empty_bar[s], STAGED ? 4 : 8);   // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;     // K blocks issued so far (over all tiles of this CTA)
      int npre = 0;
      {
        // first ring of W tiles of the first tile: weights never depend on the previous kernel
        const int t0 = blockIdx.x;
        if (t0 < n_tiles && !p.w_dynamic) {
          const int n0 = (t0 % tiles_n) * BN;
          npre = min(STAGES, nkb);
          for (int s = 0; s < npre; ++s) {
            ptx::mbar_expect_tx(&full_bar[s], L::kStageBytes);
            ptx::tma_load_3d(&tmB, &full_bar[s], sB + s * L::kBTileBytes, s * BK, n0, 0);
          }
        }
      }
      pdl_wait();
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tile_n = t % tiles_n, tile_m = t / tiles_n;
        const int m0 = tile_m * BM, n0 = tile_n * BN;
        int b0 = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          b0 = m0 / p.HW;
          const int rem = m0 - b0 * p.HW;
          h0 = rem / p.W;
          w0 = rem - h0 * p.W;
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const bool pre = it < npre;
          if (!pre) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          }
          if (p.conv) {
            const int tap = kb / p.cblocks;
            const int cb = kb - tap * p.cblocks;
            const int i = tap / p.kw;
            const int j = tap - i * p.kw;
            ptx::tma_load_4d(&tmA, &full_bar[stage], sA + stage * kATileBytes, cb * BK, w0 + j * p.dil_w - p.pad_w,
                             h0 + i * p.dil_h - p.pad_h, b0);
          } else {
            ptx::tma_load_3d(&tmA, &full_bar[stage], sA + stage * kATileBytes, kb * BK, m0, 0);
          }
          if (!pre) ptx::tma_load_3d(&tmB, &full_bar[stage], sB + stage * L::kBTileBytes, kb * BK, n0, 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int i = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
        const int acc = i & 1;
        const uint32_t use = static_cast<uint32_t>(i >> 1);
        // wait until the epilogue has drained this accumulator buffer (first use: passes on the fresh barrier)
        ptx::mbar_wait(&tmem_empty_bar[acc], (use & 1) ^ 1);
        ptx::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + stage * kATileBytes);
          const uint32_t b_addr = ptx::smem_u32(sB + stage * L::kBTileBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = ptx::umma_desc_k_sw128(a_addr + k * 32);
            const uint64_t db = ptx::umma_desc_k_sw128(b_addr + k * 32);
            ptx::umma_bf16_ss(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::umma_commit(&tmem_full_bar[acc]);
      }
      pdl_trigger();
    }
  } else {
    // ===================== epilogue (warps 2..5 [6..9] -> TMEM lane quadrants 2,3,0,1) =====================
    const int quad = warp & 3;
    const int etid = (int)threadIdx.x - 64;
    constexpr int kEpiThreads = persist_threads<STAGED>() - 64;
    constexpr int kChunksPerWarp = (BN / 32) / (kEpiThreads / 128);      // column chunks of 32 per epilogue warp
    const int c_first = ((warp - 2) >> 2) * kChunksPerWarp;
    pdl_wait();
    int i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      const int acc = i & 1;
      const uint32_t par = static_cast<uint32_t>(i >> 1) & 1u;
      const int tile_n = t % tiles_n, tile_m = t / tiles_n;
      const int m0 = tile_m * BM, n0 = tile_n * BN;
      const uint32_t tmem_q = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quad * 32) << 16);
      if (STAGED && p.fast_epi) {
        float* strip = strips + (size_t)quad * 32 * (BN + 4);
        epilogue_strip<BN>(p, strip, tmem_q, &tmem_full_bar[acc], (long long)m0 + quad * 32, n0, 0, lane, par,
                           &tmem_empty_bar[acc]);
      } else {
        const long long m = (long long)m0 + quad * 32 + lane;
        const bool row_ok = m < p.M;
        if (p.residual && row_ok) {
          const char* rp = reinterpret_cast<const char*>(p.residual + m * p.ld_res + n0);
#pragma unroll
          for (int l = 0; l < BN * 4 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + l * 128));
        }
        const bool stage_b = p.bias != nullptr;
        const long long mlast = min((long long)m0 + BM - 1, (long long)p.M - 1);
        const bool stage_rb = p.rowbias != nullptr && (m0 / p.rows_per_group) == (mlast / p.rows_per_group);
        float* sb = s_bias + acc * BN;
        float* srb = s_rowb + acc * BN;
        if (stage_b || stage_rb) {
          for (int c = etid; c < BN; c += kEpiThreads) {
            const bool ok = n0 + c < p.N;
            if (stage_b) sb[c] = ok ? __ldg(p.bias + n0 + c) : 0.f;
            if (stage_rb) srb[c] = ok ? __ldg(p.rowbias + (m0 / p.rows_per_group) * p.ld_rowbias + n0 + c) : 0.f;
          }
        }
        // also orders the epilogue warps tile by tile: buffer `acc` of sb / srb is rewritten two tiles later, after
        // every warp has passed the barrier of the tile in between
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        ptx::mbar_wait(&tmem_full_bar[acc], par);
        ptx::tcgen05_fence_after();
#pragma unroll 1
        for (int c = c_first; c < c_first + kChunksPerWarp; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(tmem_q + c * 32, v);
          ptx::tmem_ld_wait();
          if (c == c_first + kChunksPerWarp - 1) {
            // last read of this accumulator buffer: hand it back before the global stores of the chunk
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
          }
          float a32[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a32[j] = __uint_as_float(v[j]) * p.alpha;
          epilogue_chunk(p, a32, m, n0 + c * 32, 0, stage_b ? sb + c * 32 : nullptr, stage_rb ? srb + c * 32 : nullptr);
        }
      }
    }
    ptx::tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// out = act(alpha * sum_s ws[s] + bias + rowbias + residual): fixed summation order s = 0..S-1 (deterministic)
struct ReduceArgs {
  const float* ws;
  int S;
  long long MN;
  int M, N;
  const float* bias;
  const float* rowbias;
  long long ld_rowbias;
  int rows_per_group;
  const float* residual;
  long long ld_res;
  float* out_f32;
  long long ld_out_f32;
  op_t* out_bf16;
  long long ld_out_bf16;
  int act;
  float alpha;
  int res_vec;   // bit 0 / 1 / 2: residual / bias / row-bias rows are 16-byte aligned (pointer and pitch)
  unsigned long long* colstats;   // optional GroupNorm column statistics of the output (splitk_reduce_stats_kernel)
  int cs_rows;
};

__global__ void __launch_bounds__(256) splitk_reduce_kernel(ReduceArgs a) {
  pdl_wait();
  bool triggered = false;
  const int n4 = a.N >> 2;
  const long long total = (long long)a.M * n4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / n4;
    const int n = (int)(i - m * n4) << 2;
    const float* w = a.ws + m * a.N + n;
    // epilogue operands first: independent of the partial sums, so one memory round trip covers everything
    float4 bia = make_float4(0.f, 0.f, 0.f, 0.f), rbi = bia, rsd = bia;
    if (a.bias) {
      const float* q = a.bias + n;
      bia = (a.res_vec & 2) ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
    if (a.rowbias) {
      const float* q = a.rowbias + (m / a.rows_per_group) * a.ld_rowbias + n;
      rbi = (a.res_vec & 4) ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
    if (a.residual) {
      const float* r = a.residual + m * a.ld_res + n;
      if (a.res_vec & 1) rsd = *reinterpret_cast<const float4*>(r);
      else rsd = make_float4(r[0], r[1], r[2], r[3]);
    }
    float4 acc = *reinterpret_cast<const float4*>(w);
    // partials are fetched four at a time (independent loads in flight), summed in the fixed order s = 1..S-1
    int s = 1;
    for (; s + 4 <= a.S; s += 4) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = *reinterpret_cast<const float4*>(w + (long long)(s + u) * a.MN);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x += t[u].x;
        acc.y += t[u].y;
        acc.z += t[u].z;
        acc.w += t[u].w;
      }
    }
    for (; s < a.S; ++s) {
      const float4 t = *reinterpret_cast<const float4*>(w + (long long)s * a.MN);
      acc.x += t.x;
      acc.y += t.y;
      acc.z += t.z;
      acc.w += t.w;
    }
    if (!triggered) {   // partial sums of the first item are in: the tail of this kernel may overlap the next launch
      pdl_trigger();
      triggered = true;
    }
    float v[4] = {acc.x * a.alpha, acc.y * a.alpha, acc.z * a.alpha, acc.w * a.alpha};
    if (a.bias) {
      v[0] += bia.x; v[1] += bia.y; v[2] += bia.z; v[3] += bia.w;
    }
    if (a.rowbias) {
      v[0] += rbi.x; v[1] += rbi.y; v[2] += rbi.z; v[3] += rbi.w;
    }
    if (a.residual) {
      v[0] += rsd.x; v[1] += rsd.y; v[2] += rsd.z; v[3] += rsd.w;
    }
    if (a.act == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = silu_f(v[k]);
    }
    if (a.out_f32) {
      float* o = a.out_f32 + m * a.ld_out_f32 + n;
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = v[k];
    }
    if (a.out_bf16) {
      op_t* o = a.out_bf16 + m * a.ld_out_bf16 + n;
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = f2op(v[k]);
    }
  }
}

// Reduce + epilogue + GroupNorm column statistics.  A block owns 32 rows x 128 columns: warp w sums / finishes rows
// 4w..4w+3 (lane = column quad; same arithmetic and order as splitk_reduce_kernel -> same output bits), the eight
// warps' per-column sums are combined through shared memory in warp order, and one fixed-point atomic per column and
// moment goes to the accumulators.  grid (ceil(N/128), ceil(M/32)).
__global__ void __launch_bounds__(256) splitk_reduce_stats_kernel(ReduceArgs a) {
  pdl_wait();
  __shared__ float s_sum[8][128];
  __shared__ float s_sq[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + lane * 4;
  const long long m_base = (long long)blockIdx.y * 32;
  const bool col_ok = n + 3 < a.N;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = cs;
  // all loads of the warp's four rows first (S partials each), then the arithmetic
  float4 bia = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.bias && col_ok) bia = (a.res_vec & 2) ? __ldg(reinterpret_cast<const float4*>(a.bias + n))
                                              : make_float4(__ldg(a.bias + n), __ldg(a.bias + n + 1), __ldg(a.bias + n + 2), __ldg(a.bias + n + 3));
  float4 acc[4], rbi[4], rsd[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long m = m_base + warp * 4 + r;
    acc[r] = rbi[r] = rsd[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!col_ok || m >= a.M) continue;
    if (a.rowbias) {
      const float* q = a.rowbias + (m / a.rows_per_group) * a.ld_rowbias + n;
      rbi[r] = (a.res_vec & 4) ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
    if (a.residual) {
      const float* q = a.residual + m * a.ld_res + n;
      rsd[r] = (a.res_vec & 1) ? *reinterpret_cast<const float4*>(q) : make_float4(q[0], q[1], q[2], q[3]);
    }
    acc[r] = *reinterpret_cast<const float4*>(a.ws + m * a.N + n);
  }
  for (int s0 = 1; s0 < a.S; s0 += 2) {
    float4 t[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long m = m_base + warp * 4 + r;
      const bool ok = col_ok && m < a.M;
#pragma unroll
      for (int u = 0; u < 2; ++u)
        t[r][u] = (ok && s0 + u < a.S) ? *reinterpret_cast<const float4*>(a.ws + (long long)(s0 + u) * a.MN + m * a.N + n)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (s0 + u < a.S) {    // fixed order s = 1..S-1, exactly as splitk_reduce_kernel
          acc[r].x += t[r][u].x; acc[r].y += t[r][u].y; acc[r].z += t[r][u].z; acc[r].w += t[r][u].w;
        }
      }
    }
  }
  pdl_trigger();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long m = m_base + warp * 4 + r;
    if (!col_ok || m >= a.M) continue;
    float v[4] = {acc[r].x * a.alpha, acc[r].y * a.alpha, acc[r].z * a.alpha, acc[r].w * a.alpha};
    if (a.bias) {
      v[0] += bia.x; v[1] += bia.y; v[2] += bia.z; v[3] += bia.w;
    }
    if (a.rowbias) {
      v[0] += rbi[r].x; v[1] += rbi[r].y; v[2] += rbi[r].z; v[3] += rbi[r].w;
    }
    if (a.residual) {
      v[0] += rsd[r].x; v[1] += rsd[r].y; v[2] += rsd[r].z; v[3] += rsd[r].w;
    }
    if (a.act == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = silu_f(v[k]);
    }
    cs.x += v[0]; cs.y += v[1]; cs.z += v[2]; cs.w += v[3];
    cq.x = fmaf(v[0], v[0], cq.x); cq.y = fmaf(v[1], v[1], cq.y); cq.z = fmaf(v[2], v[2], cq.z); cq.w = fmaf(v[3], v[3], cq.w);
    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + m * a.ld_out_f32 + n) = make_float4(v[0], v[1], v[2], v[3]);
    if (a.out_bf16) {
      op_t* o = a.out_bf16 + m * a.ld_out_bf16 + n;
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = f2op(v[k]);
    }
  }
  *reinterpret_cast<float4*>(&s_sum[warp][lane * 4]) = cs;
  *reinterpret_cast<float4*>(&s_sq[warp][lane * 4]) = cq;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x, col = blockIdx.x * 128 + c;
    if (col < a.N && m_base < a.M) {
      float su = s_sum[0][c], sq = s_sq[0][c];
#pragma unroll
      for (int w = 1; w < 8; ++w) {
        su += s_sum[w][c];
        sq += s_sq[w][c];
      }
      colstats_add(a.colstats, m_base / a.cs_rows, a.N, col, su, sq);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// bf16 tensor map; dims/strides innermost first; strides in BYTES for dims 1..rank-1
int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(AE_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(AE_EINVAL, "TMA base pointer not 16-byte aligned");
  for (int i = 0; i < rank - 1; ++i)
    if (gs[i] % 16 != 0) return fail(AE_EINVAL, "TMA stride %d = %llu bytes not a multiple of 16", i, (unsigned long long)gs[i]);
  CUresult r = enc(tm, AE_TMAP_OPERAND_TYPE, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AE_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return AE_OK;
}

struct ConvBox {
  int Wb, Hb, Bb;
};

bool conv_box(int B, int H, int W, ConvBox* bx) {
  if (W <= 0 || H <= 0) return false;
  if (W >= BM) {
    if (W % BM != 0) return false;
    *bx = {BM, 1, 1};
    return true;
  }
  if (BM % W != 0) return false;
  const int rows = BM / W;  // image rows per tile
  if (H >= rows) {
    if (H % rows != 0) return false;
    *bx = {W, rows, 1};
    return true;
  }
  if (rows % H != 0) return false;
  *bx = {W, H, rows / H};
  return true;
}

template <int BN, int STAGES, bool STAGED>
int launch_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& p, long long tiles_m, cudaStream_t st) {
  using L = PersistLayout<BN, STAGES, STAGED>;
  static bool attr_set = false;
  static int n_sm = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_persistent_kernel<BN, STAGES, STAGED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return fail(AE_ECUDA, "cudaFuncSetAttribute(persistent smem=%d): %s", L::kTotal, cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    attr_set = true;
  }
  if (g_skip_mask & 1) return AE_OK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const long long n_tiles = tiles_m * tiles_n;
  const unsigned grid = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
  cudaError_t e = launch_kernel_early(gemm_persistent_kernel<BN, STAGES, STAGED>, dim3(grid),
                                      dim3(persist_threads<STAGED>()), (size_t)L::kTotal, st, tmA, tmB, p, tiles_n,
                                      (int)n_tiles);
  if (e != cudaSuccess) return fail(AE_ECUDA, "ae_gemm persistent launch: %s", cudaGetErrorString(e));
  return launched("ae_gemm(persistent)");
}

template <int BN, int STAGES>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& p, int gz, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  static bool attr_set = false;
  constexpr int kMaxDyn = L::kTotal;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (e != cudaSuccess) return fail(AE_ECUDA, "cudaFuncSetAttribute(smem=%d): %s", kMaxDyn, cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t dyn_smem = (size_t)L::kTotal;
  if (g_skip_mask & 1) return AE_OK;
  const unsigned tm = (unsigned)((p.M + BM - 1) / BM), tn = (unsigned)((p.N + BN - 1) / BN);
  dim3 grid = p.m_in_x ? dim3(tm, tn, gz) : dim3(tn, tm, gz);
  cudaError_t e = launch_kernel_early(gemm_tcgen05_kernel<BN, STAGES>, grid, dim3(kThreads), dyn_smem, st, tmA, tmB, p);
  if (e != cudaSuccess) return fail(AE_ECUDA, "ae_gemm launch: %s", cudaGetErrorString(e));
  return launched("ae_gemm");
}

}  // namespace

// operand-type tensor map (128-byte swizzle) for other translation units (attn_tc.cu)
int make_operand_tmap(void* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  return make_tmap(reinterpret_cast<CUtensorMap*>(tm), base, rank, dims, strides_bytes, box);
}
}  // namespace aedit

using namespace aedit;

extern "C" void ae_set_pdl(int mode) { g_use_pdl = (mode == 1 || mode == 2) ? mode : 0; }
extern "C" void ae_set_launch_priority(int prio) { g_launch_priority = prio; }
extern "C" void ae_set_skip_mask(int mask) { g_skip_mask = mask; }
extern "C" void ae_set_pdl_extra(int mask) { g_pdl_extra = mask; }
extern "C" int ae_greatest_priority(void) {
  int least = 0, greatest = 0;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return greatest;   // numerically lowest value = highest priority (0 if the device has a single level)
}

static thread_local int g_splitk_ctas = 148;
static thread_local int g_fast_epi = 1;
static thread_local long long g_persist_min_tiles = 296;   // two waves of 148 SMs; 0 = never (see ae_set_persistent_min_tiles)
extern "C" void ae_set_persistent_min_tiles(int tiles) { g_persist_min_tiles = tiles; }
static thread_local double g_reduce_us = 2.5, g_reduce_bw = 3.0e6;   // tile model: cost of the split-K reduce launch / its bytes per us
extern "C" void ae_set_tile_model_reduce(int launch_ns, int bytes_per_us) {
  g_reduce_us = launch_ns * 1e-3;
  g_reduce_bw = bytes_per_us;
}

static thread_local int g_shallow_kb = 0;
extern "C" void ae_set_shallow_kblocks(int kb) { g_shallow_kb = kb; }
static thread_local int g_shared_sm = 0;
extern "C" void ae_set_shared_sm(int on) { g_shared_sm = on ? 1 : 0; }
static thread_local int g_tile_model = 1;
extern "C" void ae_set_tile_model(int on) { g_tile_model = on ? 1 : 0; }
extern "C" void ae_set_fast_epilogue(int mode) { g_fast_epi = (mode == 0 || mode == 2) ? mode : 1; }
extern "C" void ae_set_splitk_ctas(int ctas) { g_splitk_ctas = ctas < 1 ? 148 : ctas; }

extern "C" int ae_gemm_conv_supported(int B, int H, int W, int C) {
  ConvBox bx;
  return (C % BK == 0 && conv_box(B, H, W, &bx)) ? 1 : 0;
}

extern "C" int ae_gemm(const ae_gemm_args* a, ae_stream stream) {
  AE_CHECK_ARG(a && a->A && a->W, "ae_gemm: null operand");
  AE_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "ae_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  AE_CHECK_ARG(a->out_f32 || a->out_bf16, "ae_gemm: no output");
  AE_CHECK_ARG(a->act >= 0 && a->act <= 3, "ae_gemm: act must be 0 (none), 1 (SiLU), 2 (GEGLU) or 3 (grouped softmax)");
  AE_CHECK_ARG(a->act != 2 || a->N % 32 == 0, "ae_gemm: GEGLU epilogue needs N %% 32 == 0 (N=%d)", a->N);
  const int batch = a->batch > 0 ? a->batch : 1;
  GemmDev p;
  p.M = a->M;
  p.N = a->N;
  p.conv = a->conv ? 1 : 0;
  p.bias = a->bias;
  p.rowbias = a->rowbias;
  p.ld_rowbias = a->ld_rowbias;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.residual = a->residual;
  p.ld_res = a->ld_res;
  p.out_f32 = a->out_f32;
  p.ld_out_f32 = a->ld_out_f32;
  p.out_bf16 = reinterpret_cast<op_t*>(a->out_bf16);
  p.ld_out_bf16 = a->ld_out_bf16;
  p.stride_out = a->stride_out;
  p.stride_res = a->stride_res;
  p.act = a->act;
  p.w_dynamic = a->w_dynamic ? 1 : 0;
  p.alpha = a->alpha == 0.0f ? 1.0f : a->alpha;
  p.H = p.W = p.HW = p.cblocks = p.kw = 1;
  p.dil_h = p.dil_w = 1;
  p.pad_h = p.pad_w = 0;
  p.split = 0;
  p.fast_epi = 0;
  p.m_in_x = (a->M + BM - 1) / BM > 65535 ? 1 : 0;
  p.sm_L = p.sm_block = p.sm_rows = 0;
  p.sm_slot = nullptr;
  p.sm_bias = nullptr;
  if (a->act == 3) {
    AE_CHECK_ARG(batch == 1 && !a->out_f32 && a->out_bf16 && !a->residual && !a->rowbias && !a->bias,
                 "ae_gemm: act 3 (grouped softmax) takes a plain bf16 output");
    AE_CHECK_ARG((a->sm_L == 8 || a->sm_L == 16 || a->sm_L == 32) && a->sm_block > 0 && a->sm_block % a->sm_L == 0 &&
                     a->N % a->sm_block == 0 && a->N % 32 == 0 && a->sm_slot && a->sm_rows > 0 && a->M % a->sm_rows == 0,
                 "ae_gemm: bad grouped-softmax geometry (L=%d block=%d N=%d rows=%d)", a->sm_L, a->sm_block, a->N, a->sm_rows);
    p.sm_L = a->sm_L;
    p.sm_block = a->sm_block;
    p.sm_rows = a->sm_rows;
    p.sm_slot = a->sm_slot;
    p.sm_bias = a->sm_bias;
  }
  p.colstats = nullptr;
  p.cs_rows = 1;
  if (a->colstats) {
    AE_CHECK_ARG(batch == 1 && a->act != 2 && a->N % 4 == 0 && a->out_f32,
                 "ae_gemm: colstats needs batch 1 and a plain fp32 output with N %% 4 == 0");
    AE_CHECK_ARG(a->cs_rows_per_sample > 0 && a->cs_rows_per_sample % 32 == 0 && a->M % a->cs_rows_per_sample == 0,
                 "ae_gemm: colstats needs rows per sample (%d) to be a multiple of 32 dividing M", a->cs_rows_per_sample);
    AE_CHECK_ARG(a->ld_out_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0,
                 "ae_gemm: colstats needs a 16-byte tileable fp32 output");
    p.colstats = reinterpret_cast<unsigned long long*>(a->colstats);
    p.cs_rows = a->cs_rows_per_sample;
  }

  CUtensorMap tmA, tmB;
  int rc;
  if (p.conv) {
    AE_CHECK_ARG(batch == 1, "ae_gemm: implicit conv does not take batch>1");
    AE_CHECK_ARG(a->kh >= 1 && a->kw >= 1 && a->C > 0, "ae_gemm: bad conv geometry");
    AE_CHECK_ARG(a->C % BK == 0, "ae_gemm: implicit conv needs C %% 64 == 0 (C=%d); use ae_im2col", a->C);
    AE_CHECK_ARG((long long)a->B * a->H * a->W_ == a->M, "ae_gemm: conv M=%d != B*H*W", a->M);
    AE_CHECK_ARG(a->K == a->kh * a->kw * a->C, "ae_gemm: conv K=%d != kh*kw*C", a->K);
    ConvBox bx;
    AE_CHECK_ARG(conv_box(a->B, a->H, a->W_, &bx), "ae_gemm: conv geometry H=%d W=%d not tileable; use ae_im2col", a->H,
                 a->W_);
    const int dh = a->dil_h > 0 ? a->dil_h : 1, dw = a->dil_w > 0 ? a->dil_w : 1;
    p.H = a->H;
    p.W = a->W_;
    p.HW = a->H * a->W_;
    p.cblocks = a->C / BK;
    p.kw = a->kw;
    p.dil_h = dh;
    p.dil_w = dw;
    p.pad_h = dh * (a->kh - 1) / 2;
    p.pad_w = dw * (a->kw - 1) / 2;
    p.num_kblocks = a->kh * a->kw * p.cblocks;
    uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W_, (uint64_t)a->H, (uint64_t)a->B};
    uint64_t str[3] = {(uint64_t)a->C * 2, (uint64_t)a->W_ * a->C * 2, (uint64_t)a->H * a->W_ * a->C * 2};
    uint32_t box[4] = {BK, (uint32_t)bx.Wb, (uint32_t)bx.Hb, (uint32_t)bx.Bb};
    rc = make_tmap(&tmA, a->A, 4, dims, str, box);
    if (rc) return rc;
  } else {
    AE_CHECK_ARG(a->lda >= a->K, "ae_gemm: lda < K");
    p.num_kblocks = (a->K + BK - 1) / BK;
    uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->M, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)a->lda * 2, (uint64_t)(batch > 1 ? a->strideA : (int64_t)a->M * a->lda) * 2};
    uint32_t box[3] = {BK, BM, 1};
    rc = make_tmap(&tmA, a->A, 3, dims, str, box);
    if (rc) return rc;
  }
  AE_CHECK_ARG(a->ldw >= a->K, "ae_gemm: ldw < K");
  p.kb_per_split = p.num_kblocks;

  // ---- tile width and K split.  Grids of at least one wave: 128-wide tiles (least shared-memory traffic per FLOP).
  //      Sub-wave grids (reverse process at batch 2) are bound by how fast ONE SM can pull its CTA's operands through
  //      L2 (~70 KB/us measured: [128x960x960] takes 9.9 / 7.7 / 6.5 us at BN = 128 / 64 / 32), so the tile width and
  //      the K split are chosen together to minimise the operand bytes per SM, charging a split for its reduce launch.
  const long long tiles_m = (a->M + BM - 1) / BM;
  const bool may_split = batch == 1 && a->act != 2 && a->act != 3 && a->splitk_ws && a->N % 4 == 0 &&
                         a->force_split != 1;
  int bn = a->force_bn;
  int S_model = 0;   // 0: no model decision (forced / large grid)
  if (bn == 0) {
    if (a->N <= 32)
      bn = 32;
    else if (a->N <= 64 || (a->N % 128 != 0 && a->N % 128 <= 64 && a->N < 512))
      bn = 64;
    else
      bn = 128;
    const long long tiles128 = tiles_m * ((a->N + 127) / 128);
    if (g_tile_model && batch == 1 && tiles128 < g_splitk_ctas && a->force_split <= 1) {
      double best = 1e30;
      const int cands[3] = {128, 64, 32};
      for (int ci = 0; ci < 3; ++ci) {
        const int c = cands[ci];
        if (a->act == 2 && c < 32) continue;
        if (c > 32 && a->N <= c / 2) continue;                       // mostly padding
        const long long t = tiles_m * ((a->N + c - 1) / c);
        const double stage_kb = 16.0 + c / 8.0;                       // A tile + W tile per K block
        int smax = 1;
        if (may_split && a->force_split == 0 && t < g_splitk_ctas) {
          smax = (int)(g_splitk_ctas / t);
          if (smax > p.num_kblocks / 4) smax = p.num_kblocks / 4;     // >= 4 K blocks per slice
          if (smax > 32) smax = 32;
          if (smax < 1) smax = 1;
        }
        for (int sp = 1; sp <= smax; sp = (sp < 4 ? sp + 1 : sp + 2)) {
          if (sp > 1 && (long long)sp * a->M * a->N * 4 > a->splitk_ws_bytes) break;
          const int kbs = (p.num_kblocks + sp - 1) / sp;
          const long long ctas_c = t * sp;
          const double waves = ctas_c <= g_splitk_ctas ? 1.0 : (double)ctas_c / g_splitk_ctas;
          // reduce pass: launch + (S partial tiles written and read back, residual, outputs) at ~3 MB/us
          const double reduce_us = sp > 1 ? g_reduce_us + (sp + 2.5) * (double)a->M * a->N * 4.0 / g_reduce_bw : 0.0;
          double cost = waves * kbs * stage_kb / 70.0          // us to stream one SM's operands
                        + (c / 128.0) * 1.0                     // epilogue
                        + reduce_us;
          if (cost < best - 1e-9) {
            best = cost;
            bn = c;
            S_model = sp;
          }
        }
      }
    }
  }
  AE_CHECK_ARG(bn == 32 || bn == 64 || bn == 128, "ae_gemm: force_bn must be 32, 64 or 128");
  const long long tiles = tiles_m * ((a->N + bn - 1) / bn);

  // ---- workspace split-K (two launches): partial tiles to an fp32 workspace, fixed-order reduce + epilogue kernel
  int S = 1;
  if (batch == 1 && a->act != 2 && a->act != 3 && a->splitk_ws && a->N % 4 == 0 && a->force_split != 1) {
    if (a->force_split > 1)
      S = a->force_split;
    else if (S_model > 0)
      S = S_model;
    else if (tiles <= 48 && p.num_kblocks >= 12) {
      S = (int)(g_splitk_ctas / tiles);
      const int max_by_k = p.num_kblocks / 6;
      if (S > max_by_k) S = max_by_k;
      if (S > 32) S = 32;
    }
    if (S > p.num_kblocks) S = p.num_kblocks;
    while (S > 1 && (long long)S * a->M * a->N * 4 > a->splitk_ws_bytes) --S;
    if (S < 2) S = 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->N, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)a->ldw * 2, (uint64_t)(batch > 1 ? a->strideW : (int64_t)a->N * a->ldw) * 2};
    uint32_t box[3] = {BK, (uint32_t)bn, 1};
    rc = make_tmap(&tmB, a->W, 3, dims, str, box);
    if (rc) return rc;
  }
  cudaStream_t st = as_stream(stream);
  int gz = batch;
  GemmDev q = p;
  if (S > 1) {
    q.split = 1;
    q.kb_per_split = (p.num_kblocks + S - 1) / S;
    S = (p.num_kblocks + q.kb_per_split - 1) / q.kb_per_split;  // no empty splits
    gz = S;
    q.bias = nullptr;
    q.rowbias = nullptr;
    q.residual = nullptr;
    q.out_bf16 = nullptr;
    q.out_f32 = a->splitk_ws;
    q.ld_out_f32 = a->N;
    q.stride_out = (long long)a->M * a->N;
    q.act = 0;
    q.alpha = 1.0f;
    q.colstats = nullptr;   // the reduce kernel sees the final values
  }
  {
    // coalesced staged epilogue: every pointer / pitch the strip touches must be 16-byte tileable (8 for bf16 rows,
    // half of that for the GEGLU output whose lanes own 2 columns)
    auto al = [](const void* ptr, uintptr_t b) { return (reinterpret_cast<uintptr_t>(ptr) & (b - 1)) == 0; };
    const bool g = q.act == 2;
    bool ok = g_fast_epi && q.act != 3 && (g ? (a->N % 32 == 0 && !q.residual && !q.rowbias) : (a->N % 4 == 0));
    ok = ok && (!q.bias || al(q.bias, 16));
    ok = ok && (!q.rowbias || (al(q.rowbias, 16) && q.ld_rowbias % 4 == 0));
    ok = ok && (!q.residual || (al(q.residual, 16) && q.ld_res % 4 == 0 && q.stride_res % 4 == 0));
    ok = ok && (!q.out_f32 || (al(q.out_f32, g ? 8 : 16) && q.ld_out_f32 % (g ? 2 : 4) == 0));
    ok = ok && (!q.out_bf16 || (al(q.out_bf16, g ? 4 : 8) && q.ld_out_bf16 % (g ? 2 : 4) == 0));
    ok = ok && q.stride_out % 4 == 0;
    // measured (profiles/r01_gemm_table_v11_{fast,slow}.log): the staged epilogue wins whenever a residual is read
    // (165 -> 111 us on [102400x384x384]+res) or a wide fp32 tile is written by a multi-wave grid; for bf16-only,
    // GEGLU and small-grid outputs its extra shared-memory round trip costs 1-2 us per tile, so those keep the
    // row-per-thread stores.  g_fast_epi: 0 never, 1 this rule, 2 whenever legal.
    const bool want = g_fast_epi == 2 || q.residual || (q.out_f32 && !g && tiles * gz > 2 * 148) || q.colstats;
    q.fast_epi = (ok && want) ? 1 : 0;
    if (q.colstats && !q.fast_epi)
      return fail(AE_EUNSUPPORTED, "ae_gemm: colstats needs the staged epilogue (16-byte tileable operands, "
                                   "ae_set_fast_epilogue != 0)");
  }
  const long long ctas = tiles * gz;
  const bool deep = a->force_stages ? (a->force_stages > 3) : (ctas <= 160);
  // g_shared_sm (ae_set_shared_sm): this launch will share the SMs with a throughput-bound grid of another stream
  // whose CTAs hold ~100 KB of shared memory each (two per SM).  A 6-stage ring (120-192 KB) would only fit after BOTH
  // of them have retired with no successor taking the slot, i.e. practically never; the deepest ring that fits beside
  // ONE such CTA (<= ~103 KB: 3 / 4 / 5 stages at BN = 128 / 64 / 32) gets the next free slot.  Same bits either way:
  // the ring depth changes the buffering, not the order of the accumulation.
  const bool shared_sm = g_shared_sm && !a->force_stages;
  // multi-wave grids with a short K loop: a 2-stage ring lets three CTAs share an SM (the CTA's fixed costs — launch,
  // TMEM allocation, pipeline ramp, epilogue — dominate its lifetime, so residency buys more than ring depth)
  // multi-wave grids: persistent CTAs with a double-buffered TMEM accumulator (gemm_persistent_kernel)
  // Where it pays (profiles/r01_gemm_table_persistent.log, B = 100): linears with a bf16-only or GEGLU output or a K
  // loop of >= 9 blocks (-10 .. -32 %).  Not the implicit convolutions (long K loops are bound by the operand bytes
  // in flight per SM, and two or three co-resident CTAs keep more rings in flight than one persistent CTA: +7 .. +55 %)
  // and not the K = 384 linears with an fp32 output, which are bound by their epilogue's HBM traffic (+5 %).
  // The automatic choice is further limited to the row-per-thread epilogue (bf16-only / GEGLU outputs — the shapes
  // that gain most), whose persistent CTA fits in ~100 KB with a 3-stage ring: a persistent CTA holds its SM for the
  // whole kernel, and with the 200 KB staged variant no CTA of the reverse-process lane could become resident anywhere
  // on the machine meanwhile (whole job: 836 -> 862 ms with every shape persistent, profiles/r01_lanes_ab6.log).
  const bool persist_auto = g_persist_min_tiles > 0 && tiles >= g_persist_min_tiles && !p.conv && !q.fast_epi &&
                            (!q.out_f32 || q.act == 2 || p.num_kblocks >= 9);
  const bool persist = S == 1 && batch == 1 && !a->force_stages && (bn == 128 || bn == 64) &&
                       (a->force_persistent > 0 || (a->force_persistent == 0 && persist_auto)) &&
                       tiles * 1ll < 2147483647ll;
  if (persist) {
    if (bn == 128)
      rc = q.fast_epi ? launch_persistent<128, 4, true>(tmA, tmB, q, tiles_m, st)
                      : launch_persistent<128, 3, false>(tmA, tmB, q, tiles_m, st);
    else
      rc = q.fast_epi ? launch_persistent<64, 6, true>(tmA, tmB, q, tiles_m, st)
                      : launch_persistent<64, 4, false>(tmA, tmB, q, tiles_m, st);
  } else if (!deep && bn == 128 && (a->force_stages == 2 || (!a->force_stages && p.num_kblocks <= g_shallow_kb)))
    rc = launch<128, 2>(tmA, tmB, q, gz, st);
  else
  switch (bn) {
    case 32:
      rc = deep ? (shared_sm ? launch<32, 5>(tmA, tmB, q, gz, st) : launch<32, 6>(tmA, tmB, q, gz, st))
                : launch<32, 3>(tmA, tmB, q, gz, st);
      break;
    case 64:
      rc = deep ? (shared_sm ? launch<64, 4>(tmA, tmB, q, gz, st) : launch<64, 6>(tmA, tmB, q, gz, st))
                : launch<64, 3>(tmA, tmB, q, gz, st);
      break;
    default:
      rc = (deep && !shared_sm) ? launch<128, 6>(tmA, tmB, q, gz, st) : launch<128, 3>(tmA, tmB, q, gz, st);
      break;
  }
  if (rc || S == 1) return rc;
  ReduceArgs r;
  r.ws = a->splitk_ws;
  r.S = S;
  r.MN = (long long)a->M * a->N;
  r.M = a->M;
  r.N = a->N;
  r.bias = p.bias;
  r.rowbias = p.rowbias;
  r.ld_rowbias = p.ld_rowbias;
  r.rows_per_group = p.rows_per_group;
  r.residual = p.residual;
  r.ld_res = p.ld_res;
  r.out_f32 = p.out_f32;
  r.ld_out_f32 = p.ld_out_f32;
  r.out_bf16 = p.out_bf16;
  r.ld_out_bf16 = p.ld_out_bf16;
  r.act = p.act;
  r.alpha = p.alpha;
  r.res_vec = ((p.residual && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0 && p.ld_res % 4 == 0) ? 1 : 0) |
              ((p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) ? 2 : 0) |
              ((p.rowbias && (reinterpret_cast<uintptr_t>(p.rowbias) & 15) == 0 && p.ld_rowbias % 4 == 0) ? 4 : 0);
  r.colstats = p.colstats;
  r.cs_rows = p.cs_rows;
  long long blocks = ceil_div64(r.MN / 4, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (g_skip_mask & 2) return AE_OK;
  cudaError_t e;
  if (r.colstats)
    e = launch_kernel_early(splitk_reduce_stats_kernel, dim3((unsigned)((a->N + 127) / 128), (unsigned)((a->M + 31) / 32)),
                            dim3(256), (size_t)0, st, r);
  else
    e = launch_kernel_early(splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, st, r);
  if (e != cudaSuccess) return fail(AE_ECUDA, "splitk_reduce launch: %s", cudaGetErrorString(e));
  return launched("ae_gemm(splitk_reduce)");
}
