// Fused multi-head attention with online softmax (K5 of SURVEY.md §2.2):
//   O = softmax(Q K^T * scale + key_bias) V        reference math: attention.py:285-323 (normal_attention),
//   additive key mask convention of models.py:204-210 (keep 0 / discard -10000).
// Used for self-attention (Tk = H*W tokens, up to 4096) and cross-attention against the frozen text K/V
// (Tk = 8..~128, shared across the batch through kv_batch_map).
//
// One CTA = 64 queries of one (batch, head); 4 warps x 16 query rows.  K/V tiles of 64 keys are double-buffered
// in shared memory with cp.async; S = QK^T and O += P V run on the warp-level tensor-core path
// (mma.sync.m16n8k16 bf16, fp32 accumulate) with P kept in registers between the two products; row statistics
// are reduced with warp shuffles inside each quad.  Every (batch, head, query) is processed identically
// regardless of batch size (batch-invariant), scores never touch HBM.
// (The tcgen05 port of this kernel — S and O accumulators in TMEM — is tracked in DESIGN.md §kernels.)
#include "common.cuh"
#include "ptx.cuh"

namespace aedit {
namespace {

constexpr int BQ = 64;
constexpr int BKV = 64;
constexpr int kAttnThreads = 128;

struct AttnArgs {
  const op_t* q;
  const op_t* k;
  const op_t* v;
  op_t* o;
  long long ld_q, ld_k, ld_v, ld_o;
  long long bs_q, bs_k, bs_v, bs_o;
  const int* kv_map;
  const float* bias;
  long long ld_bias;
  int Tq, Tk;
  float scale_log2;  // scale * log2(e)
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int LDS = DP + 8;  // padded row (elements): conflict-free ldmatrix
  // K/V ring depth.  Deeper rings (3-4 stages) were measured: no gain at B=2, -8 % at B=16 (the loop is issue-bound,
  // not load-latency-bound, and the extra shared memory costs occupancy) -> double buffering.
  static constexpr int NS = 2;
  static constexpr int kSmemBytes = (BQ + 2 * NS * BKV) * LDS * 2;
};

template <int D>
__device__ __forceinline__ void load_tile(op_t* dst, const op_t* src, long long ld, int row0, int rows,
                                          int tid) {
  using C = AttnCfg<D>;
  constexpr int CH = C::DP / 8;  // 16-byte chunks per padded row
  for (int i = tid; i < 64 * CH; i += kAttnThreads) {
    const int r = i / CH, c = i % CH;
    const bool ok = (row0 + r < rows) && (c * 8 < D);
    const op_t* g = ok ? src + (long long)(row0 + r) * ld + c * 8 : src;
    ptx::cp_async_16(ptx::smem_u32(dst + r * C::LDS + c * 8), g, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads) attention_kernel(AttnArgs a) {
  pdl_wait();
  using C = AttnCfg<D>;
  constexpr int DP = C::DP;
  constexpr int LDS = C::LDS;
  constexpr int KS = DP / 16;  // k-steps over the head dim for QK^T
  constexpr int NB = DP / 8;   // output n-blocks (head-dim columns / 8)
  extern __shared__ __align__(16) uint8_t smem_attn[];
  op_t* sQ = reinterpret_cast<op_t*>(smem_attn);
  op_t* sK = sQ + BQ * LDS;
  constexpr int NS = C::NS;
  op_t* sV = sK + NS * BKV * LDS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bkv = a.kv_map ? a.kv_map[b] : b;
  const op_t* qp = a.q + (long long)b * a.bs_q + (long long)head * D;
  const op_t* kp = a.k + (long long)bkv * a.bs_k + (long long)head * D;
  const op_t* vp = a.v + (long long)bkv * a.bs_v + (long long)head * D;
  const float* biasp = a.bias ? a.bias + (long long)bkv * a.ld_bias : nullptr;

  constexpr int tile0 = 0;
  const int ntiles = (a.Tk + BKV - 1) / BKV;
  // prologue: Q and the first NS-1 key/value tiles, one commit group per tile (empty groups keep the count uniform)
  load_tile<D>(sQ, qp, a.ld_q, q0, a.Tq, tid);
#pragma unroll
  for (int t = 0; t < NS - 1; ++t) {
    if (t < ntiles) {
      load_tile<D>(sK + t * BKV * LDS, kp, a.ld_k, (tile0 + t) * BKV, a.Tk, tid);
      load_tile<D>(sV + t * BKV * LDS, vp, a.ld_v, (tile0 + t) * BKV, a.Tk, tid);
    }
    ptx::cp_async_commit();
  }

  float o_acc[NB][4];
#pragma unroll
  for (int n = 0; n < NB; ++n)
#pragma unroll
    for (int k = 0; k < 4; ++k) o_acc[n][k] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[KS][4];

  for (int it = 0; it < ntiles; ++it) {
    const int buf = it % NS;
    {
      // refill the slot consumed in the previous iteration (protected by that iteration's trailing barrier)
      const int nt = it + NS - 1;
      if (nt < ntiles) {
        const int nb = nt % NS;
        load_tile<D>(sK + nb * BKV * LDS, kp, a.ld_k, (tile0 + nt) * BKV, a.Tk, tid);
        load_tile<D>(sV + nb * BKV * LDS, vp, a.ld_v, (tile0 + nt) * BKV, a.Tk, tid);
      }
      ptx::cp_async_commit();
      ptx::cp_async_wait<NS - 1>();   // tile `it` (and Q) have landed
    }
    __syncthreads();
    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ptx::ldmatrix_x4(qf[ks], ptx::smem_u32(sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8));
    }
    const op_t* tK = sK + buf * BKV * LDS;
    const op_t* tV = sV + buf * BKV * LDS;

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k) s[n][k] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key blocks
        uint32_t bf[4];
        const int mi = lane >> 3;
        const int key = np * 16 + (mi >> 1) * 8 + (lane & 7);
        const int dcol = ks * 16 + (mi & 1) * 8;
        ptx::ldmatrix_x4(bf, ptx::smem_u32(tK + key * LDS + dcol));
        ptx::mma_m16n8k16_bf16(s[2 * np], qf[ks], bf[0], bf[1]);
        ptx::mma_m16n8k16_bf16(s[2 * np + 1], qf[ks], bf[2], bf[3]);
      }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16).  The raw scores stay unscaled: scale*log2(e) is
    //      folded into the exponent FMA.  Key-range masking / additive bias only run on tiles that need them.
    const int key0 = (tile0 + it) * BKV;
    const bool full_tile = (key0 + BKV <= a.Tk);
    if (biasp != nullptr || !full_tile) {
      const float inv = 1.0f / a.scale_log2;   // bring the bias into raw-score units
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int key = key0 + n * 8 + t4 * 2 + (k & 1);
          float val = s[n][k];
          if (key < a.Tk) {
            if (biasp) val += biasp[key] * (1.4426950408889634f * inv);
          } else {
            val = -INFINITY;
          }
          s[n][k] = val;
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
    bool changed = false;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);
      changed |= (m_new != m_run[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : fast_exp2((m_run[r] - m_new) * a.scale_log2);
      m_run[r] = m_new;
      msc[r] = m_new * a.scale_log2;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A-operand fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float p0 = fast_exp2(fmaf(s[n][0], a.scale_log2, -msc[0]));
      const float p1 = fast_exp2(fmaf(s[n][1], a.scale_log2, -msc[0]));
      const float p2 = fast_exp2(fmaf(s[n][2], a.scale_log2, -msc[1]));
      const float p3 = fast_exp2(fmaf(s[n][3], a.scale_log2, -msc[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      const op2_t h01 = ff2op2(p0, p1);
      const op2_t h23 = ff2op2(p2, p3);
      pf[n >> 1][(n & 1) * 2 + 0] = *reinterpret_cast<const uint32_t*>(&h01);
      pf[n >> 1][(n & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
    if (__any_sync(0xffffffffu, changed)) {   // running max unchanged for the whole warp -> corr == 1, skip the rescale
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        o_acc[n][0] *= corr[0];
        o_acc[n][1] *= corr[0];
        o_acc[n][2] *= corr[1];
        o_acc[n][3] *= corr[1];
      }
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < NB / 2; ++np) {
        uint32_t bf[4];
        const int mi = lane >> 3;
        const int key = kk * 16 + (mi & 1) * 8 + (lane & 7);
        const int dcol = np * 16 + (mi >> 1) * 8;
        ptx::ldmatrix_x4_trans(bf, ptx::smem_u32(tV + key * LDS + dcol));
        ptx::mma_m16n8k16_bf16(o_acc[2 * np], pf[kk], bf[0], bf[1]);
        ptx::mma_m16n8k16_bf16(o_acc[2 * np + 1], pf[kk], bf[2], bf[3]);
      }
    }
    __syncthreads();  // all warps done with this buffer before the next prefetch overwrites it
  }

  pdl_trigger();   // key/value loop done: the normalise + store tail may overlap the next launch
  // ---- normalise and store (bf16 pairs)
  const int r0 = q0 + warp * 16 + g;
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  op_t* op = a.o + (long long)b * a.bs_o + (long long)head * D;
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    const int col = n * 8 + t4 * 2;
    if (col < D) {
      if (r0 < a.Tq)
        *reinterpret_cast<op2_t*>(op + (long long)r0 * a.ld_o + col) =
            ff2op2(o_acc[n][0] * inv0, o_acc[n][1] * inv0);
      if (r0 + 8 < a.Tq)
        *reinterpret_cast<op2_t*>(op + (long long)(r0 + 8) * a.ld_o + col) =
            ff2op2(o_acc[n][2] * inv1, o_acc[n][3] * inv1);
    }
  }
}

template <int D>
int launch_attn(const AttnArgs& a, int B, int heads, cudaStream_t st) {
  using C = AttnCfg<D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return fail(AE_ECUDA, "attention smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((a.Tq + BQ - 1) / BQ, heads, B);
  launch_kernel_family(8, attention_kernel<D>, dim3(grid), dim3(kAttnThreads), (size_t)(C::kSmemBytes), st, a);
  return launched("ae_attention");
}

}  // namespace
}  // namespace aedit

namespace aedit {
// attn_tc.cu: the tcgen05 / TMEM / TMA attention path.  Returns 1 if it took the call (*rc = status), else 0.
int attention_tc_try(const void* q, int64_t ld_q, int64_t q_bs, const void* k, int64_t ld_k, int64_t k_bs, const void* v,
                     int64_t ld_v, int64_t v_bs, const int32_t* kv_map, const float* key_bias, int64_t ld_bias, int B,
                     int Bkv, int heads, int d, int Tq, int Tk, float scale, void* out, int64_t ld_o, int64_t o_bs,
                     cudaStream_t st, int* rc);
}  // namespace aedit

using namespace aedit;

static int attention_impl(const void* q, int64_t ld_q, int64_t q_bs, const void* k, int64_t ld_k, int64_t k_bs,
                          const void* v, int64_t ld_v, int64_t v_bs, const int32_t* kv_batch_map, const float* key_bias,
                          int64_t ld_bias, int B, int heads, int d, int Tq, int Tk, float scale, void* out, int64_t ld_o,
                          int64_t o_bs, ae_stream stream) {

  AE_CHECK_ARG(q && k && v && out && B > 0 && heads > 0 && Tq > 0 && Tk > 0, "ae_attention: bad argument");
  AE_CHECK_ARG(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_o % 2 == 0,
               "ae_attention: row strides must keep 16-byte alignment");
  AttnArgs a;
  a.q = reinterpret_cast<const op_t*>(q);
  a.k = reinterpret_cast<const op_t*>(k);
  a.v = reinterpret_cast<const op_t*>(v);
  a.o = reinterpret_cast<op_t*>(out);
  a.ld_q = ld_q;
  a.ld_k = ld_k;
  a.ld_v = ld_v;
  a.ld_o = ld_o;
  a.bs_q = q_bs;
  a.bs_k = k_bs;
  a.bs_v = v_bs;
  a.bs_o = o_bs;
  a.kv_map = kv_batch_map;
  a.bias = key_bias;
  a.ld_bias = ld_bias;
  a.Tq = Tq;
  a.Tk = Tk;
  a.scale_log2 = scale * 1.4426950408889634f;
  cudaStream_t st = as_stream(stream);
  if (g_skip_mask & 32) return AE_OK;
  {
    int rc = AE_OK;      // long sequences: tensor-core (tcgen05) kernel; everything else: the mma.sync kernel below
    if (attention_tc_try(q, ld_q, q_bs, k, ld_k, k_bs, v, ld_v, v_bs, kv_batch_map, key_bias, ld_bias, B, B, heads, d, Tq, Tk,
                         scale, out, ld_o, o_bs, st, &rc))
      return rc;
  }
  switch (d) {
    case 32: return launch_attn<32>(a, B, heads, st);
    case 40: return launch_attn<40>(a, B, heads, st);
    case 48: return launch_attn<48>(a, B, heads, st);
    case 64: return launch_attn<64>(a, B, heads, st);
    case 72: return launch_attn<72>(a, B, heads, st);
    case 80: return launch_attn<80>(a, B, heads, st);
    case 96: return launch_attn<96>(a, B, heads, st);
    case 120: return launch_attn<120>(a, B, heads, st);
    case 128: return launch_attn<128>(a, B, heads, st);
    case 160: return launch_attn<160>(a, B, heads, st);
    default:
      return fail(AE_EUNSUPPORTED, "ae_attention: head dim %d not instantiated (32,40,48,64,72,80,96,120,128,160)", d);
  }
}

extern "C" int ae_attention(const void* q, int64_t ld_q, int64_t q_bs, const void* k, int64_t ld_k, int64_t k_bs,
                            const void* v, int64_t ld_v, int64_t v_bs, const int32_t* kv_batch_map, const float* key_bias,
                            int64_t ld_bias, int B, int heads, int d, int Tq, int Tk, float scale, void* out, int64_t ld_o,
                            int64_t o_bs, ae_stream stream) {
  return attention_impl(q, ld_q, q_bs, k, ld_k, k_bs, v, ld_v, v_bs, kv_batch_map, key_bias, ld_bias, B, heads, d, Tq, Tk,
                        scale, out, ld_o, o_bs, stream);
}
