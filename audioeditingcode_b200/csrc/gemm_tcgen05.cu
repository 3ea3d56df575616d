// tcgen05 / TMA GEMM and implicit-GEMM convolution (K1/K2/K5-linear of SURVEY.md §2.2).
//
//   D[M,N] = alpha * A[M,K] · W[N,K]^T (+bias)(+rowbias)(+residual) -> act -> {f32, bf16} outputs
//
// Replaces every cuDNN/cuBLAS call the reference's U-Net evaluation lands in (conv3x3 / conv1x1 / Linear of
// diffusers ResnetBlock2D + Transformer2DModel driven from code/models.py:293-388; in-tree twins
// code/audioldm/latent_diffusion/openaimodel.py:213-244, attention.py:220-323).
//
// Tile: 128 (M) x BN (N) x 64 (K) bf16, fp32 accumulators in TMEM.  One CTA per output tile, 6 warps:
//   warp 0  TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, STAGES-deep mbarrier ring)
//   warp 1  TMEM allocator + single-thread tcgen05.mma issuer (4 UMMAs of K=16 per stage) + tcgen05.commit
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused bias / time-embedding row bias /
//              residual / SiLU, direct vectorised global stores (each thread owns one output row segment)
// Implicit convolution: the A operand is a channels-last image [B,H,W,C]; K-block kb = (tap, 64-channel slab);
// the producer issues a 4-D TMA box {64 ch, Wb, Hb, Bb} at (c0, w0+dw, h0+dh, b0) — out-of-bounds rows/cols are
// zero-filled by TMA, which is exactly the convolution's zero padding.  The 128 tile rows are the flattened
// (b,h,w) positions m0..m0+127, so the epilogue is identical to the plain GEMM.
// Determinism / batch invariance: no split-K, no atomics — a sample's reduction order never depends on the
// batch it is launched with (needed for the bit-exact replay invariant).
#include <cuda.h>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace aedit {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kATileBytes = BM * BK * 2;  // 16 KiB

struct GemmDev {
  int M, N;
  int num_kblocks;
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
  __nv_bfloat16* out_bf16;
  long long ld_out_bf16;
  long long stride_out, stride_res;
  int act;
  float alpha;
};

template <int BN>
struct SmemLayout {
  static constexpr int kBTileBytes = BN * BK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (BN <= 32) ? 8 : (BN <= 64 ? 8 : 6);
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024 manual alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  using L = SmemLayout<BN>;
  constexpr int STAGES = L::kStages;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms (TMA and UMMA both derive the XOR from address bits)
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kATileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x;
  const int tile_n = blockIdx.y;
  const int z = blockIdx.z;
  const int m0 = tile_m * BM;
  const int n0 = tile_n * BN;

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
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

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
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_kblocks; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
        if (p.conv) {
          const int tap = kb / p.cblocks;
          const int cb = kb - tap * p.cblocks;
          const int i = tap / p.kw;
          const int j = tap - i * p.kw;
          ptx::tma_load_4d(&tmA, &full_bar[stage], sA + stage * kATileBytes, cb * BK, w0 + j * p.dil_w - p.pad_w,
                           h0 + i * p.dil_h - p.pad_h, b0);
        } else {
          ptx::tma_load_3d(&tmA, &full_bar[stage], sA + stage * kATileBytes, kb * BK, m0, z);
        }
        ptx::tma_load_3d(&tmB, &full_bar[stage], sB + stage * L::kBTileBytes, kb * BK, n0, z);
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
      for (int kb = 0; kb < p.num_kblocks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tcgen05_fence_after();
        const uint32_t a_addr = ptx::smem_u32(sA + stage * kATileBytes);
        const uint32_t b_addr = ptx::smem_u32(sB + stage * L::kBTileBytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128 B swizzle atom
          const uint64_t da = ptx::umma_desc_k_sw128(a_addr + k * 32);
          const uint64_t db = ptx::umma_desc_k_sw128(b_addr + k * 32);
          ptx::umma_bf16_ss(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      ptx::umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const long long m = (long long)m0 + row;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
    const bool row_ok = m < p.M;
    const float* rb = nullptr;
    if (p.rowbias && row_ok) rb = p.rowbias + (m / p.rows_per_group) * p.ld_rowbias;
    const float* res = p.residual ? p.residual + (long long)z * p.stride_res + m * p.ld_res : nullptr;
    float* of = p.out_f32 ? p.out_f32 + (long long)z * p.stride_out + m * p.ld_out_f32 : nullptr;
    __nv_bfloat16* ob = p.out_bf16 ? p.out_bf16 + (long long)z * p.stride_out + m * p.ld_out_bf16 : nullptr;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
      ptx::tmem_ld_wait();
      const int nbase = n0 + c * 32;
      if (!row_ok || nbase >= p.N) continue;
      const int nvalid = min(32, p.N - nbase);
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]) * p.alpha;
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) acc[j] += __ldg(p.bias + nbase + j);
      }
      if (rb) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) acc[j] += __ldg(rb + nbase + j);
      }
      if (res) {
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
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = silu_f(acc[j]);
      }
      if (of) {
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(of + nbase) & 15) == 0)) {
          float4* o4 = reinterpret_cast<float4*>(of + nbase);
#pragma unroll
          for (int j = 0; j < 8; ++j) o4[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) of[nbase + j] = acc[j];
        }
      }
      if (ob) {
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(ob + nbase) & 15) == 0)) {
          uint4* o4 = reinterpret_cast<uint4*>(ob + nbase);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[8 * j + 0], acc[8 * j + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(acc[8 * j + 2], acc[8 * j + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[8 * j + 4], acc[8 * j + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(acc[8 * j + 6], acc[8 * j + 7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2);
            pk.w = *reinterpret_cast<uint32_t*>(&h3);
            o4[j] = pk;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) ob[nbase + j] = __float2bfloat16_rn(acc[j]);
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
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
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

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& p, int batch, cudaStream_t st) {
  using L = SmemLayout<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return fail(AE_ECUDA, "cudaFuncSetAttribute(smem=%d): %s", L::kTotal, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, batch);
  gemm_tcgen05_kernel<BN><<<grid, kThreads, L::kTotal, st>>>(tmA, tmB, p);
  return launched("ae_gemm");
}

}  // namespace
}  // namespace aedit

using namespace aedit;

extern "C" int ae_gemm_conv_supported(int B, int H, int W, int C) {
  ConvBox bx;
  return (C % BK == 0 && conv_box(B, H, W, &bx)) ? 1 : 0;
}

extern "C" int ae_gemm(const ae_gemm_args* a, ae_stream stream) {
  AE_CHECK_ARG(a && a->A && a->W, "ae_gemm: null operand");
  AE_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "ae_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  AE_CHECK_ARG(a->out_f32 || a->out_bf16, "ae_gemm: no output");
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
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.ld_out_bf16 = a->ld_out_bf16;
  p.stride_out = a->stride_out;
  p.stride_res = a->stride_res;
  p.act = a->act;
  p.alpha = a->alpha == 0.0f ? 1.0f : a->alpha;
  p.H = p.W = p.HW = p.cblocks = p.kw = 1;
  p.dil_h = p.dil_w = 1;
  p.pad_h = p.pad_w = 0;

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

  // tile width: wide tiles when the grid already fills the machine, narrower ones to spread weight streaming
  int bn = a->force_bn;
  if (bn == 0) {
    const long long tiles_m = (a->M + BM - 1) / BM;
    bn = 128;
    if (a->N <= 32)
      bn = 32;
    else if (a->N <= 64)
      bn = 64;
    else {
      const long long t128 = tiles_m * ((a->N + 127) / 128) * batch;
      const long long t64 = tiles_m * ((a->N + 63) / 64) * batch;
      if (t128 < 120) bn = (t64 < 120) ? 32 : 64;
    }
  }
  AE_CHECK_ARG(bn == 32 || bn == 64 || bn == 128, "ae_gemm: force_bn must be 32, 64 or 128");
  {
    uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->N, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)a->ldw * 2, (uint64_t)(batch > 1 ? a->strideW : (int64_t)a->N * a->ldw) * 2};
    uint32_t box[3] = {BK, (uint32_t)bn, 1};
    rc = make_tmap(&tmB, a->W, 3, dims, str, box);
    if (rc) return rc;
  }
  cudaStream_t st = as_stream(stream);
  switch (bn) {
    case 32:
      return launch<32>(tmA, tmB, p, batch, st);
    case 64:
      return launch<64>(tmA, tmB, p, batch, st);
    default:
      return launch<128>(tmA, tmB, p, batch, st);
  }
}
