// Fused multi-head attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA): the Blackwell-native path of
// ae_attention for the long sequences of the U-Net (self-attention over T = H*W tokens, code/audioldm/latent_diffusion/
// attention.py:285-323: softmax(Q K^T * scale + bias) V).  The mma.sync kernel of attn_kernels.cu stays as the path for
// short sequences (T < 128) and odd layouts.
//
// One CTA owns NT tiles of 128 queries of one (batch, head) and walks the keys in blocks of BKEYS:
//   warp 4*NT*SP   TMA producer: Q tile(s) once, then a 2-stage ring of K / V blocks (cp.async.bulk.tensor, 128-byte
//                  swizzle; the head's d columns are addressed through a 4-D tensor map {d, heads, T, batch}, so columns
//                  beyond d and rows beyond T are zero-filled by the TMA unit = the padding to 64 / 128 columns)
//   warp 4*NT*SP+1 TMEM allocator + single-thread MMA issuer:
//                    S = Q K^T      tcgen05.mma kind::f16, A = Q (K-major), B = K block (K-major), D = S in TMEM (fp32)
//                    O_blk = P V    A = P (K-major, written to shared memory by the softmax warps),
//                                   B = V block exactly as TMA delivers it = MN-major operand (ptx::umma_desc_mn_sw128)
//   warps 0..4*NT*SP-1  softmax: query row i = TMEM lane i is owned by SP threads (SP = 2: half of the key columns each): S row,
//                  scale / key bias, block maximum, P = exp2(s - m_ref) -> operand type -> shared memory in the swizzled
//                  K-major layout, row sum.  The output accumulates IN TMEM across the key blocks (tcgen05.mma with the
//                  accumulate flag): TMEM reads cost 64 B/clk, so S is read exactly once per block and O only when it must
//                  be rescaled — the reference maximum m_ref of a row is raised only when the block maximum exceeds it by
//                  more than 8 (in log2 units; P then stays <= 256, exact enough in fp16 / bf16 with fp32 accumulation),
//                  in which case the warp rescales its O rows in place (tcgen05.ld / tcgen05.st) before the next P.V.
// NT = 2 query tiles per CTA (8 softmax warps = 2 per SM sub-partition) when the grid fills the machine anyway: MUFU, TMEM
// reads and FMA work only overlap ACROSS warps, and the tiles share each K / V block; NT = 1 (double-buffered S) for small
// grids (reverse process at batch 2).  A row's arithmetic does not depend on NT: same bits.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace aedit {
namespace {

constexpr int kQT = 128;   // queries per tile = TMEM lanes

struct AttnTcArgs {
  const int* kv_map;       // [B] K/V batch of query batch b, or null
  const float* key_bias;   // [Bkv, Tk] additive (fp32), or null
  long long ld_bias;
  int d, d16, Tq, Tk, nblk;
  float scale_log2;        // scale * log2(e)
  int need_mask;           // key bias present or a partial last key block
  op_t* out;
  long long ld_o, bs_o;
};

template <int DP, int BKEYS, int NT, int SP>
struct TcCfg {
  static constexpr int kStages = 2;
  static constexpr int kDBlocks = DP / 64;                 // 64-column d blocks (one TMA box each)
  static constexpr int kQTileBytes = kQT * DP * 2;
  static constexpr int kKVBlockBytes = BKEYS * DP * 2;     // one K (or V) block
  static constexpr int kPTileBytes = kQT * BKEYS * 2;
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + NT * kQTileBytes;
  static constexpr int kOffV = kOffK + kStages * kKVBlockBytes;
  static constexpr int kOffP = kOffV + kStages * kKVBlockBytes;
  static constexpr int kOffBias = kOffP + NT * kPTileBytes;
  static constexpr int kOffX = kOffBias + NT * 4 * BKEYS * 4;        // row-statistics exchange between the SP column halves
  static constexpr int kOffBar = kOffX + NT * 3 * 2 * kQT * 4;
  static constexpr int kTotal = kOffBar + 256 + 1024;      // + manual 1024-byte alignment slack
  static constexpr int kThreads = (4 * NT * SP + 2) * 32;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(NT * BKEYS <= 256 && (NT == 1 ? 2 * BKEYS <= 256 : true) && NT * DP <= 256, "TMEM column plan");
  static constexpr int kColS(int buf) { return buf * BKEYS; }    // S tiles (NT = 1: two buffers) in columns [0, 256)
  static constexpr int kColO(int t) { return 256 + t * DP; }     // O tiles in columns [256, 512)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DP, int BKEYS, int NT, int SP>
__global__ void __launch_bounds__(TcCfg<DP, BKEYS, NT, SP>::kThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnTcArgs a) {
  using C = TcCfg<DP, BKEYS, NT, SP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [2]
  uint64_t* kv_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;        // [4]  NT > 1: per tile; NT = 1: per S buffer
  uint64_t* p_full = bars + 9;        // [4]  per tile
  uint64_t* o_full = bars + 13;       // [4]  per tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (kQT * NT);
  constexpr int kSoftmaxWarps = 4 * NT * SP;

  if (threadIdx.x == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(kv_full + i, 1);
      ptx::mbar_init(kv_empty + i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(s_full + i, 1);
      ptx::mbar_init(p_full + i, kQT * SP);
      ptx::mbar_init(o_full + i, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kSoftmaxWarps + 1) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  if (warp == kSoftmaxWarps && lane == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmV);
  }
  pdl_trigger();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nblk = a.nblk;
  pdl_wait();

  if (warp == kSoftmaxWarps) {
    // ================================================================= TMA producer
    if (lane == 0) {
      const int kvb = a.kv_map ? a.kv_map[b] : b;
      ptx::mbar_expect_tx(q_full, NT * C::kQTileBytes);
      for (int t = 0; t < NT; ++t)
        for (int db = 0; db < C::kDBlocks; ++db)
          ptx::tma_load_4d(&tmQ, q_full, smem + C::kOffQ + t * C::kQTileBytes + db * (kQT * 128), db * 64, head,
                           q0 + t * kQT, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        if (j >= 2) ptx::mbar_wait(kv_empty + st, ((j >> 1) - 1) & 1);
        ptx::mbar_expect_tx(kv_full + st, 2 * C::kKVBlockBytes);
        for (int db = 0; db < C::kDBlocks; ++db) {
          ptx::tma_load_4d(&tmK, kv_full + st, smem + C::kOffK + st * C::kKVBlockBytes + db * (BKEYS * 128), db * 64, head,
                           j * BKEYS, kvb);
          ptx::tma_load_4d(&tmV, kv_full + st, smem + C::kOffV + st * C::kKVBlockBytes + db * (BKEYS * 128), db * 64, head,
                           j * BKEYS, kvb);
        }
      }
    }
  } else if (warp == kSoftmaxWarps + 1) {
    // ================================================================= MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(kQT, BKEYS);
      const uint32_t idesc_o = ptx::umma_idesc_f32_b_mn(kQT, a.d16);
      const int nks = a.d16 / 16;
      const uint32_t q_addr = ptx::smem_u32(smem + C::kOffQ), k_addr = ptx::smem_u32(smem + C::kOffK);
      const uint32_t v_addr = ptx::smem_u32(smem + C::kOffV), p_addr = ptx::smem_u32(smem + C::kOffP);
      auto issue_s = [&](int t, int st, int sbuf) {        // S_t = Q_t . K_st^T  -> TMEM columns kColS(sbuf)
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t off_a = (ks >> 2) * (kQT * 128) + (ks & 3) * 32;
          const uint32_t off_b = (ks >> 2) * (BKEYS * 128) + (ks & 3) * 32;
          ptx::umma_bf16_ss(tmem_base + C::kColS(sbuf), ptx::umma_desc_k_sw128(q_addr + t * C::kQTileBytes + off_a),
                            ptx::umma_desc_k_sw128(k_addr + st * C::kKVBlockBytes + off_b), idesc_s, ks > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int st, bool first) {      // O_t (+)= P_t . V_st  -> TMEM columns kColO(t)
        for (int kk = 0; kk < BKEYS / 16; ++kk) {
          const uint32_t off_a = (kk >> 2) * (kQT * 128) + (kk & 3) * 32;
          ptx::umma_bf16_ss(tmem_base + C::kColO(t), ptx::umma_desc_k_sw128(p_addr + t * C::kPTileBytes + off_a),
                            ptx::umma_desc_mn_sw128(v_addr + st * C::kKVBlockBytes + kk * 2048, BKEYS * 128), idesc_o,
                            (!first || kk > 0) ? 1u : 0u);
        }
      };
      ptx::mbar_wait(q_full, 0);
      ptx::tcgen05_fence_after();
      if constexpr (NT == 1) {
        ptx::mbar_wait(kv_full, 0);
        ptx::tcgen05_fence_after();
        issue_s(0, 0, 0);
        ptx::umma_commit(s_full);
        for (int j = 0; j < nblk; ++j) {
          if (j + 1 < nblk) {                               // S(j+1) into the other buffer while softmax(j) runs
            const int st = (j + 1) & 1;
            ptx::mbar_wait(kv_full + st, ((j + 1) >> 1) & 1);
            ptx::tcgen05_fence_after();
            issue_s(0, st, (j + 1) & 1);
            ptx::umma_commit(s_full + ((j + 1) & 1));
          }
          ptx::mbar_wait(p_full, j & 1);
          ptx::tcgen05_fence_after();
          issue_pv(0, j & 1, j == 0);
          ptx::umma_commit(o_full);
          ptx::umma_commit(kv_empty + (j & 1));
        }
      } else {
        for (int j = 0; j <= nblk; ++j) {
          if (j < nblk) {
            ptx::mbar_wait(kv_full + (j & 1), (j >> 1) & 1);
            ptx::tcgen05_fence_after();
          }
          for (int t = 0; t < NT; ++t) {
            if (j > 0) {                                    // P_t(j-1) is in shared memory (and S_t(j-1) fully read)
              ptx::mbar_wait(p_full + t, (j - 1) & 1);
              ptx::tcgen05_fence_after();
              issue_pv(t, (j - 1) & 1, j == 1);
              ptx::umma_commit(o_full + t);
            }
            if (j < nblk) {
              issue_s(t, j & 1, t);
              ptx::umma_commit(s_full + t);
            }
          }
          if (j > 0) ptx::umma_commit(kv_empty + ((j - 1) & 1));
        }
      }
    }
  } else {
    // ================================================================= softmax warps (thread = query row)
    // SP = 2: a query row is shared by TWO threads (warps w and w + 4 of the tile own the same 32 TMEM lanes), each
    // handling half of the key columns of every block and half of the output columns: the per-block critical path of a
    // row (TMEM read -> max -> exp2 -> pack -> store) halves, at the price of one maximum exchange per block.
    const int t = warp / (4 * SP);                          // tile
    const int half = (warp >> 2) % SP;                      // column half owned by this warp
    const int row = (warp & 3) * 32 + lane;                 // TMEM lane / row of the tile
    const int q = q0 + t * kQT + row;
    constexpr int KC = BKEYS / SP;                          // key columns per thread and block
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float* sbias = reinterpret_cast<float*>(smem + C::kOffBias) + (t * 4 * SP + (warp & 3) * SP + half) * KC;  // per warp
    float* xch = reinterpret_cast<float*>(smem + C::kOffX) + t * (3 * 2 * kQT);       // [2 parities + final][2 halves][128]
    uint8_t* p_row = smem + C::kOffP + t * C::kPTileBytes + row * 128;
    const int kvb = a.kv_map ? a.kv_map[b] : b;
    const float* bias_row = a.key_bias ? a.key_bias + (long long)kvb * a.ld_bias : nullptr;
    const int oc0 = half * (a.d16 / SP), oc1 = oc0 + a.d16 / SP;                      // output columns of this thread
    auto tile_sync = [&]() {
      if (SP > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "r"(kQT * SP) : "memory");
    };
    float m_ref = -1.0e30f, l_run = 0.f;       // m_ref: the maximum the row's P values / O accumulator refer to (log2 units)
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleGap = 8.0f;        // raise m_ref only when a block maximum exceeds it by more than 2^8

#pragma unroll 1
    for (int j = 0; j < nblk; ++j) {
      const int sbuf = (NT == 1) ? (j & 1) : t;
      const uint32_t s_addr = lane_addr + C::kColS(sbuf);
      if (a.need_mask) {                                    // additive key bias / out-of-range keys of this block
        __syncwarp();
        for (int i = lane; i < KC; i += 32) {
          const int key = j * BKEYS + half * KC + i;
          sbias[i] = key < a.Tk ? (bias_row ? bias_row[key] * kLog2e : 0.f) : -1.0e30f;
        }
        __syncwarp();
      }
      ptx::mbar_wait(s_full + ((NT == 1) ? (j & 1) : t), (NT == 1) ? ((j >> 1) & 1) : (j & 1));
      ptx::tcgen05_fence_after();
      // ---- the S row, read from TMEM exactly once
      uint32_t v[KC];
#pragma unroll
      for (int c = 0; c < KC / 32; ++c)
        ptx::tmem_ld_32x32b_x32(s_addr + half * KC + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
      ptx::tmem_ld_wait();
      float mx4[4] = {-1.0e30f, -1.0e30f, -1.0e30f, -1.0e30f};     // four independent chains (one warp per sub-partition: ILP)
      if (a.need_mask) {
#pragma unroll
        for (int i = 0; i < KC; ++i) {
          const float x = fmaf(__uint_as_float(v[i]), a.scale_log2, sbias[i]);
          v[i] = __float_as_uint(x);
          mx4[i & 3] = fmaxf(mx4[i & 3], x);
        }
      } else {
#pragma unroll
        for (int i = 0; i < KC; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if (!a.need_mask) mx *= a.scale_log2;                 // scale > 0: max commutes with the scaling
      if (SP > 1) {                                         // block maximum of the whole row: exchange the two halves
        float* x2 = xch + (j & 1) * (2 * kQT);
        x2[half * kQT + row] = mx;
        tile_sync();
        mx = fmaxf(mx, x2[(half ^ 1) * kQT + row]);
      }
      // ---- reference maximum: first block sets it; later blocks raise it (and rescale O, l) only past the gap
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool raise = mx > m_ref + kRescaleGap;
        if (__any_sync(0xffffffffu, raise)) {
          ptx::mbar_wait(o_full + t, (j - 1) & 1);          // P.V of block j-1 has landed: O is stable
          ptx::tcgen05_fence_after();
          const float m_new = raise ? mx : m_ref;
          const float alpha = ex2f(m_ref - m_new);
#pragma unroll 1
          for (int c = oc0; c < oc1; c += 8) {              // rare path: 8 columns at a time keeps the register peak low
            uint32_t o[8];
            ptx::tmem_ld_32x32b_x8(lane_addr + C::kColO(t) + c, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st_32x32b_x8(lane_addr + C::kColO(t) + c, o);
          }
          ptx::tmem_st_wait();
          l_run *= alpha;
          m_ref = m_new;
        }
      }
      // ---- P = exp2(x - m_ref) -> operand type -> shared memory (K-major, 128-byte swizzle), row sum
      if (j > 0) {
        // single query tile: S(j) was issued ahead of P.V(j-1) (double-buffered S), so that product may still be READING
        // the P tile this block is about to overwrite — wait for it.  With two tiles S_t(j) is issued after P.V_t(j-1) and
        // tensor-core work completes in order, so seeing S_t(j) already implies it and the wait returns at once; it is
        // kept so that every phase of o_full is observed before the next commit arrives on it (synccheck: missing wait)
        ptx::mbar_wait(o_full + t, (j - 1) & 1);
      }
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c8l = 0; c8l < KC / 8; ++c8l) {
        const int c8 = half * (KC / 8) + c8l;               // 16-byte chunk of the row's BKEYS keys
        float pv[8];
        if (a.need_mask) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pv[i] = ex2f(__uint_as_float(v[c8l * 8 + i]) - m_ref);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) pv[i] = ex2f(fmaf(__uint_as_float(v[c8l * 8 + i]), a.scale_log2, -m_ref));
        }
        ls4[c8l & 3] += ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
        op2_t h0 = ff2op2(pv[0], pv[1]), h1 = ff2op2(pv[2], pv[3]), h2 = ff2op2(pv[4], pv[5]), h3 = ff2op2(pv[6], pv[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        uint8_t* blk = p_row + (c8 >> 3) * (kQT * 128);               // 64-key block of the P tile
        *reinterpret_cast<uint4*>(blk + (((c8 & 7) ^ (row & 7)) << 4)) = pk;
      }
      l_run += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      ptx::fence_proxy_async_smem();          // generic-proxy stores of P -> visible to the tensor core's async proxy
      ptx::tcgen05_fence_before();            // this thread's TMEM accesses (S read, O rescale) are complete
      ptx::mbar_arrive(p_full + t);
    }
    // ---- normalise and store (this thread's output columns)
    if (SP > 1) {                                           // row sum = the two halves' partial sums (same m_ref)
      float* x2 = xch + 2 * (2 * kQT);
      x2[half * kQT + row] = l_run;
      tile_sync();
      l_run += x2[(half ^ 1) * kQT + row];
    }
    ptx::mbar_wait(o_full + t, (nblk - 1) & 1);
    ptx::tcgen05_fence_after();
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    op_t* dst = a.out + (long long)b * a.bs_o + (long long)q * a.ld_o + (long long)head * a.d;
#pragma unroll 1
    for (int c = oc0; c < oc1; c += 8) {
      uint32_t o[8];
      ptx::tmem_ld_32x32b_x8(lane_addr + C::kColO(t) + c, o);
      ptx::tmem_ld_wait();
      if (q < a.Tq && c < a.d) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[i]) * inv;
        op2_t h0 = ff2op2(f[0], f[1]), h1 = ff2op2(f[2], f[3]), h2 = ff2op2(f[4], f[5]), h3 = ff2op2(f[6], f[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + c) = pk;
      }
    }
    ptx::tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == kSoftmaxWarps + 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int DP, int BKEYS, int NT, int SP>
int launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcArgs& a, int B, int heads,
              cudaStream_t st) {
  using C = TcCfg<DP, BKEYS, NT, SP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<DP, BKEYS, NT, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal);
    if (e != cudaSuccess) return fail(AE_ECUDA, "attn_tc smem attribute (%d bytes): %s", C::kTotal, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((a.Tq + kQT * NT - 1) / (kQT * NT), heads, B);
  launch_kernel_family(8, attn_tc_kernel<DP, BKEYS, NT, SP>, grid, dim3(C::kThreads), (size_t)C::kTotal, st, tq, tk, tv, a);
  return launched("ae_attention(tcgen05)");
}

}  // namespace

static thread_local int g_attn_tc = 1;     // 1: tcgen05 path where applicable (default); 0: always the mma.sync kernel

// Returns 1 if the tcgen05 path took the call (rc holds its status), 0 if the caller should use the mma.sync kernel.
int attention_tc_try(const void* q, int64_t ld_q, int64_t q_bs, const void* k, int64_t ld_k, int64_t k_bs, const void* v,
                     int64_t ld_v, int64_t v_bs, const int32_t* kv_map, const float* key_bias, int64_t ld_bias, int B,
                     int Bkv, int heads, int d, int Tq, int Tk, float scale, void* out, int64_t ld_o, int64_t o_bs,
                     cudaStream_t st, int* rc) {
  if (!g_attn_tc || kv_map != nullptr || d % 8 != 0 || d > 128 || d < 16 || Tq < kQT || Tk < 128) return 0;
  if (heads > 65535 || B > 65535 || ld_o % 8 != 0 || o_bs % 8 != 0) return 0;
  const int64_t strides[6] = {ld_q, q_bs, ld_k, k_bs, ld_v, v_bs};
  for (int64_t s : strides)
    if (s % 8 != 0) return 0;                                  // TMA strides are multiples of 16 bytes
  if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return 0;
  const int DP = d <= 64 ? 64 : 128;
  const long long tiles = (long long)B * heads * ((Tq + kQT - 1) / kQT);
  // Two query tiles per CTA (two softmax warps per SM sub-partition, shared K / V blocks) when the grid fills the machine
  // anyway, one (double-buffered S) for small grids.  Measured and rejected (profiles/r02_attn_bench_v4.log): four tiles
  // per CTA with 64-key blocks (16 softmax warps under a 96-register cap) — 6-28 % SLOWER: the kernel is bound by the
  // per-block barrier round trips (S ready -> softmax -> P ready -> P.V -> next S), which twice as many, half as large
  // key blocks double.
  // g_attn_tc: 1 auto; 2 / 3 / 4 / 5 force (NT 1, SP 2) / (NT 2, SP 1) / (NT 1, SP 1) / (NT 2, SP 2) — A/B measurements.
  // Auto (profiles/r02_attn_bench_v5.log): a row split over two threads halves the per-block critical path of the softmax
  // warps and wins while the grid is below ~3 waves of CTAs; beyond that two query tiles per CTA (shared K / V blocks)
  // win — with the row split on top for d <= 64 (16 softmax warps; for d > 64 the extra warps gain nothing).
  int NT = tiles >= 3 * 148 ? 2 : 1, SP = (NT == 1 || DP == 64) ? 2 : 1;
  if (g_attn_tc == 2) { NT = 1; SP = 2; }
  if (g_attn_tc == 3) { NT = 2; SP = 1; }
  if (g_attn_tc == 4) { NT = 1; SP = 1; }
  if (g_attn_tc == 5) { NT = 2; SP = 2; }
  const int BKEYS = DP == 64 ? 128 : 64;
  CUtensorMap tq, tk, tv;
  auto mk = [&](CUtensorMap* tm, const void* base, int64_t ld, int64_t bs, int T, int nb, int rows) {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)T, (uint64_t)nb};
    const uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)(nb > 1 ? bs : (int64_t)T * ld) * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)rows, 1};
    return make_operand_tmap(tm, base, 4, dims, str, box);
  };
  *rc = mk(&tq, q, ld_q, q_bs, Tq, B, kQT);
  if (*rc == AE_OK) *rc = mk(&tk, k, ld_k, k_bs, Tk, Bkv, BKEYS);
  if (*rc == AE_OK) *rc = mk(&tv, v, ld_v, v_bs, Tk, Bkv, BKEYS);
  if (*rc != AE_OK) return 1;
  AttnTcArgs a;
  a.kv_map = kv_map;
  a.key_bias = key_bias;
  a.ld_bias = ld_bias;
  a.d = d;
  a.d16 = (d + 15) / 16 * 16;
  a.Tq = Tq;
  a.Tk = Tk;
  a.nblk = (Tk + BKEYS - 1) / BKEYS;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.need_mask = (key_bias != nullptr || Tk % BKEYS != 0) ? 1 : 0;
  a.out = reinterpret_cast<op_t*>(out);
  a.ld_o = ld_o;
  a.bs_o = o_bs;
  if (DP == 64) {
    if (NT == 2 && SP == 2) *rc = launch_tc<64, 128, 2, 2>(tq, tk, tv, a, B, heads, st);
    else if (NT == 2) *rc = launch_tc<64, 128, 2, 1>(tq, tk, tv, a, B, heads, st);
    else if (SP == 2) *rc = launch_tc<64, 128, 1, 2>(tq, tk, tv, a, B, heads, st);
    else *rc = launch_tc<64, 128, 1, 1>(tq, tk, tv, a, B, heads, st);
  } else {
    if (NT == 2 && SP == 2) *rc = launch_tc<128, 64, 2, 2>(tq, tk, tv, a, B, heads, st);
    else if (NT == 2) *rc = launch_tc<128, 64, 2, 1>(tq, tk, tv, a, B, heads, st);
    else if (SP == 2) *rc = launch_tc<128, 64, 1, 2>(tq, tk, tv, a, B, heads, st);
    else *rc = launch_tc<128, 64, 1, 1>(tq, tk, tv, a, B, heads, st);
  }
  return 1;
}

}  // namespace aedit

extern "C" void ae_set_attention_tc(int on) { aedit::g_attn_tc = (on >= 0 && on <= 5) ? on : 1; }
