// Unsupervised principal-direction editing on the device (reference code/pc_drift.py):
//   ae_pc_perturb        pc_drift.py:41-42, :64-80   input = xt + amount * eigvecs * sqrt(alpha_bar_t), laid out as the
//                                                    2n-row CFG batch (uncond rows, cond rows) of one U-Net launch
//   ae_pc_subspace_step  pc_drift.py:148-185         Ab = x0_pred*mask - x0_ref; per-direction norm; normalise;
//                                                    re-orthonormalise (torch.linalg.qr + sign rule + sort); correlation
//                                                    with the previous iterate; next perturbation = const * eigvecs
//   ae_pc_apply_drift    pc_drift.py:232-278         shift along sum_k amount*sqrt(eigval_k)*eigvec_k and re-compose x_{t-1}
//
// Re-orthonormalisation = CholeskyQR2 instead of LAPACK's Householder sweep over a [D x n] matrix: ONE pass over the
// n direction vectors accumulates the Gram matrix (n(n+1)/2 sums) and the masked norms, a single warp factors the
// n x n Gram matrix, a second pass applies the n x n transform; repeated once on the result (orthogonality error
// ~ eps instead of ~ cond(G)*eps).  Every pass is a grid-wide streaming kernel over n*D floats (1 MiB at n = 8 for a
// 10 s clip) with per-CTA partial sums combined in a FIXED order (deterministic, no floating-point atomics).
// The thin Q of a QR factorisation is unique up to the sign of each column; torch.linalg.qr (LAPACK geqrf/orgqr)
// fixes that sign through the Householder rule R_kk = -sign(alpha_k)*||x_k||, alpha_k being the k-th element of the
// k-th partially reduced column.  The reference's results (signs of the returned directions, its `swap` rule on
// prod(diag R), its correlation traces) depend on it, so the rule is reproduced exactly: the whole Householder
// recursion lives in span{e_0..e_{n-1}, v_0..v_{n-1}}, i.e. it can be carried out on 2n coefficients given the Gram
// matrix and the first n rows of the matrix — which is what pc_finalize_kernel does, in double precision.
#include "common.cuh"

namespace aedit {
namespace {

constexpr int kPcMaxN = 16;
constexpr int kPcThreads = 256;
constexpr int kPcMaxCtas = 148;

__host__ __device__ constexpr int pc_nacc(int n) { return n * (n + 1) / 2 + n; }

struct PcWs {            // layout of the caller's workspace (all offsets in floats / doubles, computed by pc_ws_layout)
  float* partials;       // [ctas, nacc]  per-CTA partial sums of the current pass
  float* q1;             // [n, n_el]     first-round Q
  double* small;         // small dense state, see indices below
};
// small[] indices
constexpr int kS_T = 0;                                   // [n*n] transform of the pending apply pass
constexpr int kS_norm = kS_T + kPcMaxN * kPcMaxN;         // [n] norms of Ab
constexpr int kS_sign = kS_norm + kPcMaxN;                // [n] Householder signs s_k (with the global swap folded in)
constexpr int kS_order = kS_sign + kPcMaxN;               // [n] sort order (as doubles)
constexpr int kS_total = kS_order + kPcMaxN;

struct PcArgs {
  const float* x0_pred;   // [n, n_el]
  const float* x0_ref;    // [n_el]
  const float* mask;      // [n_el] or null (all ones)
  const float* prev;      // [n, n_el] or null
  float* eig_out;         // [n, n_el]
  float* eig_scaled;      // [n, n_el] or null
  float* norms_out;       // [n]
  float* corr_out;        // [n] or null
  float cnst;
  int n;
  int64_t n_el;
  int ctas;
  PcWs ws;
};

// value of direction i at element d for the two source kinds
__device__ __forceinline__ float pc_masked_ab(const PcArgs& a, int i, int64_t d, float m, float ref, float* sq) {
  const float ab = __fsub_rn(__fmul_rn(a.x0_pred[(int64_t)i * a.n_el + d], m), ref);   // pc_drift.py:148-149
  *sq = (m != 0.0f) ? ab * ab : 0.0f;                                                     // norm over mask.bool() (:158)
  return ab * m;                                                                          // (... ) * mask (:160)
}

template <int NMAX>
__device__ __forceinline__ void pc_block_reduce_store(float (&acc)[pc_nacc(NMAX)], int nacc_rt, float* dst) {
  __shared__ float s_red[kPcThreads / 32][pc_nacc(NMAX)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < pc_nacc(NMAX); ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nacc_rt; k += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) v += s_red[w][k];
    dst[k] = v;
  }
}

// accumulator index of pair (i <= j) for a COMPILE-TIME layout of NMAX: row-major upper triangle, then the NMAX norms
template <int NMAX>
__device__ __forceinline__ constexpr int pc_pair(int i, int j) { return i * NMAX - i * (i - 1) / 2 + (j - i); }

// Pass A: Gram matrix + masked norms of the masked differences (SRC = 0) or of plain vectors (SRC = 1, no norms).
template <int NMAX, int SRC>
__global__ void __launch_bounds__(kPcThreads) pc_gram_kernel(PcArgs a, const float* __restrict__ src) {
  pdl_trigger();
  pdl_wait();
  float acc[pc_nacc(NMAX)];
#pragma unroll
  for (int k = 0; k < pc_nacc(NMAX); ++k) acc[k] = 0.f;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < a.n_el; d += (int64_t)gridDim.x * blockDim.x) {
    float v[NMAX];
    if (SRC == 0) {
      const float m = a.mask ? a.mask[d] : 1.0f;
      const float ref = a.x0_ref[d];
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        float sq = 0.f;
        v[i] = (i < a.n) ? pc_masked_ab(a, i, d, m, ref, &sq) : 0.f;
        acc[NMAX * (NMAX + 1) / 2 + i] += sq;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NMAX; ++i) v[i] = (i < a.n) ? src[(int64_t)i * a.n_el + d] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
#pragma unroll
      for (int j = i; j < NMAX; ++j) acc[pc_pair<NMAX>(i, j)] += v[i] * v[j];
  }
  pc_block_reduce_store<NMAX>(acc, pc_nacc(NMAX), a.ws.partials + (int64_t)blockIdx.x * pc_nacc(NMAX));
}

// Pass B: out_k = sum_i T[k][i] * src_i  (SRC = 0: src_i = masked difference, out -> ws.q1, accumulates the Gram matrix
// of the outputs for the second round; SRC = 1: src_i = ws.q1, out -> eig_out / eig_scaled, accumulates <prev_k, out_k>)
template <int NMAX, int SRC>
__global__ void __launch_bounds__(kPcThreads) pc_apply_kernel(PcArgs a, int final_pass) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sT[NMAX * NMAX];
  for (int k = threadIdx.x; k < NMAX * NMAX; k += blockDim.x) {
    const int r = k / NMAX, c = k % NMAX;
    sT[k] = (r < a.n && c < a.n) ? (float)a.ws.small[kS_T + r * kPcMaxN + c] : 0.f;
  }
  __syncthreads();
  float acc[pc_nacc(NMAX)];
#pragma unroll
  for (int k = 0; k < pc_nacc(NMAX); ++k) acc[k] = 0.f;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < a.n_el; d += (int64_t)gridDim.x * blockDim.x) {
    float v[NMAX], o[NMAX];
    if (SRC == 0) {
      const float m = a.mask ? a.mask[d] : 1.0f;
      const float ref = a.x0_ref[d];
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        float sq;
        v[i] = (i < a.n) ? pc_masked_ab(a, i, d, m, ref, &sq) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NMAX; ++i) v[i] = (i < a.n) ? a.ws.q1[(int64_t)i * a.n_el + d] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NMAX; ++i) s += sT[k * NMAX + i] * v[i];
      o[k] = s;
    }
    if (!final_pass) {
#pragma unroll
      for (int k = 0; k < NMAX; ++k)
        if (k < a.n) a.ws.q1[(int64_t)k * a.n_el + d] = o[k];
#pragma unroll
      for (int i = 0; i < NMAX; ++i)
#pragma unroll
        for (int j = i; j < NMAX; ++j) acc[pc_pair<NMAX>(i, j)] += o[i] * o[j];
    } else {
#pragma unroll
      for (int k = 0; k < NMAX; ++k) {
        if (k < a.n) {
          const int64_t idx = (int64_t)k * a.n_el + d;
          if (a.prev) acc[k] += a.prev[idx] * o[k];                        // pc_drift.py:181-182 (diag of prev @ new^T)
          a.eig_out[idx] = o[k];
          if (a.eig_scaled) a.eig_scaled[idx] = __fmul_rn(o[k], a.cnst);   // pc_drift.py:193
        }
      }
    }
  }
  pc_block_reduce_store<NMAX>(acc, final_pass ? NMAX : pc_nacc(NMAX), a.ws.partials + (int64_t)blockIdx.x * pc_nacc(NMAX));
}

// Small dense algebra between the passes: one warp, double precision.  stage 0: after the first Gram pass (norms,
// Cholesky, Householder sign emulation, sort order, transform of round 1); stage 1: after the second Gram pass (final
// transform with signs / permutation); stage 2: correlation sums.
template <int NMAX>
__global__ void __launch_bounds__(32) pc_finalize_kernel(PcArgs a, int stage) {
  pdl_trigger();
  pdl_wait();
  constexpr int NACC = pc_nacc(NMAX);
  constexpr int N2 = 2 * NMAX;
  __shared__ double sum[NACC];
  __shared__ double G[NMAX][NMAX], R[NMAX][NMAX], Ri[NMAX][NMAX];
  __shared__ double M[N2][N2], Cw[NMAX][N2], cx[N2], cv[N2], Mcv[N2];
  const int lane = threadIdx.x;
  const int n = a.n;
  const int nsum = (stage == 2) ? NMAX : NACC;
  for (int k = lane; k < nsum; k += 32) {
    double s = 0.0;
    for (int c = 0; c < a.ctas; ++c) s += (double)a.ws.partials[(int64_t)c * NACC + k];   // fixed order
    sum[k] = s;
  }
  __syncwarp();
  double* S = a.ws.small;
  if (stage == 2) {
    if (a.corr_out)
      for (int k = lane; k < n; k += 32) a.corr_out[k] = (float)sum[k];
    return;
  }
  auto pair = [](int i, int j) { return i <= j ? pc_pair<NMAX>(i, j) : pc_pair<NMAX>(j, i); };
  if (stage == 0) {
    for (int k = lane; k < n; k += 32) {
      const double nr = sqrt(sum[NMAX * (NMAX + 1) / 2 + k]);
      S[kS_norm + k] = nr;
      a.norms_out[k] = (float)nr;                                         // norm_of_Ab (pc_drift.py:158 / :176)
    }
    __syncwarp();
    if (n == 1) {                                                         // pc_drift.py:175-177: no QR, no sign rule
      if (lane == 0) {
        S[kS_T] = 1.0 / S[kS_norm];
        S[kS_sign] = 1.0;
        S[kS_order] = 0.0;
      }
      return;
    }
    // Gram matrix of the normalised masked vectors V_i = w_i / norm_i
    for (int k = lane; k < n * n; k += 32) {
      const int i = k / n, j = k % n;
      G[i][j] = sum[pair(i, j)] / (S[kS_norm + i] * S[kS_norm + j]);
    }
    __syncwarp();
    // ---- Householder sign rule on 2n coefficients: basis [e_0..e_{n-1} | V_0..V_{n-1}], metric M
    for (int k = lane; k < N2 * N2; k += 32) M[k / N2][k % N2] = 0.0;
    __syncwarp();
    for (int k = lane; k < n; k += 32) M[k][k] = 1.0;
    for (int k = lane; k < n * n; k += 32) {
      const int r = k / n, i = k % n;            // At[r][i] = V_i[r], the first n rows of the [D x n] matrix
      const float m = a.mask ? a.mask[r] : 1.0f;
      float sq;
      const double val = (double)pc_masked_ab(a, i, r, m, a.x0_ref[r], &sq) / S[kS_norm + i];
      M[r][NMAX + i] = val;
      M[NMAX + i][r] = val;
      M[NMAX + r][NMAX + i] = G[r][i];
    }
    for (int k = lane; k < NMAX * N2; k += 32) Cw[k / N2][k % N2] = 0.0;
    __syncwarp();
    for (int k = lane; k < n; k += 32) Cw[k][NMAX + k] = 1.0;
    __syncwarp();
    double sign_prod = 1.0;
    for (int k = 0; k < n; ++k) {
      // x = W_k with the rows above k removed: cx = c_k - sum_{r<k} (e_r . W_k) e_r
      for (int q = lane; q < N2; q += 32) {
        double v = Cw[k][q];
        if (q < k) {
          double er = 0.0;
          for (int p = 0; p < N2; ++p) er += M[q][p] * Cw[k][p];
          v -= er;
        }
        cx[q] = v;
      }
      __syncwarp();
      double nx2 = 0.0, alpha = 0.0;
      {                                   // every lane computes the two scalars redundantly (n is tiny)
        for (int q = 0; q < N2; ++q) {
          double mq = 0.0;
          for (int p = 0; p < N2; ++p) mq += M[q][p] * cx[p];
          nx2 += cx[q] * mq;
          if (q == k) alpha = mq;
        }
      }
      const double nx = sqrt(fmax(nx2, 0.0));
      const double beta = (alpha >= 0.0) ? -nx : nx;          // LAPACK slarfg: beta = -sign(norm, alpha)
      if (lane == 0) S[kS_sign + k] = (beta >= 0.0) ? 1.0 : -1.0;
      sign_prod *= (beta >= 0.0) ? 1.0 : -1.0;
      for (int q = lane; q < N2; q += 32) cv[q] = cx[q] - ((q == k) ? beta : 0.0);
      __syncwarp();
      for (int q = lane; q < N2; q += 32) {
        double mq = 0.0;
        for (int p = 0; p < N2; ++p) mq += M[q][p] * cv[p];
        Mcv[q] = mq;
      }
      __syncwarp();
      double vv = 0.0;
      for (int q = 0; q < N2; ++q) vv += cv[q] * Mcv[q];
      for (int j = k + 1; j < n; ++j) {
        double proj = 0.0;
        for (int q = 0; q < N2; ++q) proj += Cw[j][q] * Mcv[q];
        const double g = (vv > 0.0) ? 2.0 * proj / vv : 0.0;
        __syncwarp();
        for (int q = lane; q < N2; q += 32) Cw[j][q] -= g * cv[q];
        __syncwarp();
      }
    }
    __syncwarp();
    if (lane == 0) {
      // swap rule (pc_drift.py:164-166): prod(diag R) < 0  ->  Q *= -1
      if (sign_prod < 0.0)
        for (int k = 0; k < n; ++k) S[kS_sign + k] = -S[kS_sign + k];
      // stable descending sort of the norms (pc_drift.py:172-173; eigenvalue = norm / const * sigma^2 is monotone in it)
      int ord[NMAX];
      for (int k = 0; k < n; ++k) ord[k] = k;
      for (int i = 1; i < n; ++i) {
        const int o = ord[i];
        int j = i - 1;
        while (j >= 0 && S[kS_norm + ord[j]] < S[kS_norm + o]) {
          ord[j + 1] = ord[j];
          --j;
        }
        ord[j + 1] = o;
      }
      for (int k = 0; k < n; ++k) S[kS_order + k] = (double)ord[k];
    }
  } else {
    for (int k = lane; k < n * n; k += 32) G[k / n][k % n] = sum[pair(k / n, k % n)];
  }
  __syncwarp();
  // Cholesky G = R^T R (upper R, positive diagonal) and R^{-1}, by lane 0 (n <= 16)
  if (lane == 0) {
    for (int j = 0; j < n; ++j) {
      for (int i = 0; i <= j; ++i) {
        double s = G[i][j];
        for (int p = 0; p < i; ++p) s -= R[p][i] * R[p][j];
        R[i][j] = (i == j) ? sqrt(fmax(s, 1e-300)) : s / R[i][i];
      }
      for (int i = j + 1; i < n; ++i) R[i][j] = 0.0;
    }
    for (int j = 0; j < n; ++j) {                       // Ri = R^{-1} (upper triangular), column by column
      for (int i = 0; i < n; ++i) Ri[i][j] = 0.0;
      Ri[j][j] = 1.0 / R[j][j];
      for (int i = j - 1; i >= 0; --i) {
        double s = 0.0;
        for (int p = i + 1; p <= j; ++p) s += R[i][p] * Ri[p][j];
        Ri[i][j] = -s / R[i][i];
      }
    }
    // q_k = sum_i Ri[i][k] * src_i
    if (stage == 0) {
      for (int k = 0; k < n; ++k)
        for (int i = 0; i < n; ++i) S[kS_T + k * kPcMaxN + i] = Ri[i][k] / S[kS_norm + i];
    } else {
      for (int k = 0; k < n; ++k) {
        const int o = (int)S[kS_order + k];
        for (int i = 0; i < n; ++i) S[kS_T + k * kPcMaxN + i] = S[kS_sign + o] * Ri[i][o];
      }
    }
  }
}

struct PerturbArgs {
  const float* xt;
  int64_t xt_stride;
  const float* eig;
  float amount, sqrt_ab;
  int mode, n;
  float* x_batch;
  float* inp_out;
  int64_t n_el;
};

__global__ void __launch_bounds__(256) pc_perturb_kernel(PerturbArgs a) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.y;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < a.n_el; d += (int64_t)gridDim.x * blockDim.x) {
    const float x = a.xt[(int64_t)i * a.xt_stride + d];
    float inp = x;
    // xt + (amount * eigvecs) * sqrt(alpha_bar_t), one rounding per op like the eager expression (pc_drift.py:42)
    if (a.eig) inp = __fadd_rn(x, __fmul_rn(__fmul_rn(a.amount, a.eig[(int64_t)i * a.n_el + d]), a.sqrt_ab));
    a.x_batch[(int64_t)i * a.n_el + d] = (a.mode == 1 || a.mode == 3) ? inp : x;              // BOTH / UNCOND (:65)
    a.x_batch[(int64_t)(a.n + i) * a.n_el + d] = (a.mode == 1 || a.mode == 2) ? inp : x;      // BOTH / TEXT   (:75)
    if (a.inp_out) a.inp_out[(int64_t)i * a.n_el + d] = inp;
  }
}

struct DriftArgs {
  const float* xt_m1;
  const float* x0_pred;
  const float* latent;
  const float* shift;     // [n_el] = sum_k amount*sqrt(eigval_k)*eigvec_k, broadcast over the batch rows
  float* out;
  float std_t, sqrt_ap, c_dir, ratio;   // eta*sqrt(var), sqrt(alpha_prev), sqrt(1-alpha_prev-std^2), sqrt(ab)/sqrt(1-ab)
  int eta_pos, use_shifted;
  int64_t n_el;
};

// pc_drift.py:232-278, one rounding per op in the reference's order
__global__ void __launch_bounds__(256) pc_apply_drift_kernel(DriftArgs a) {
  pdl_trigger();
  pdl_wait();
  const int64_t row = (int64_t)blockIdx.y * a.n_el;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < a.n_el; d += (int64_t)gridDim.x * blockDim.x) {
    const float sh = a.shift[d];
    const float x0p = a.x0_pred[row + d];
    const float x0d = __fadd_rn(x0p, sh);                                             // :236
    float xm1 = a.xt_m1[row + d];
    const float noise = a.eta_pos ? __fmul_rn(a.std_t, a.latent[row + d]) : 0.f;
    if (a.eta_pos) xm1 = __fsub_rn(xm1, noise);                                        // :251-252
    const float dir = __fsub_rn(xm1, __fmul_rn(a.sqrt_ap, x0p));                       // :255
    float pe = __fdiv_rn(dir, a.c_dir);                                                // :256
    if (a.use_shifted) pe = __fsub_rn(pe, __fmul_rn(a.ratio, sh));                     // :258-259
    const float dir2 = __fmul_rn(a.c_dir, pe);                                         // :263
    float o = __fadd_rn(__fmul_rn(a.sqrt_ap, x0d), dir2);                              // :266
    if (a.eta_pos) o = __fadd_rn(o, noise);                                            // :268-269
    a.out[row + d] = o;
  }
}

template <int NMAX>
int pc_step_launch(PcArgs a, cudaStream_t st) {
  const dim3 grid(a.ctas), blk(kPcThreads);
  launch_kernel(pc_gram_kernel<NMAX, 0>, grid, blk, (size_t)0, st, a, (const float*)nullptr);
  if (int rc = launched("pc_gram")) return rc;
  launch_kernel(pc_finalize_kernel<NMAX>, dim3(1), dim3(32), (size_t)0, st, a, 0);
  if (int rc = launched("pc_finalize")) return rc;
  if (a.n == 1) {
    launch_kernel(pc_apply_kernel<NMAX, 0>, grid, blk, (size_t)0, st, a, 1);
    if (int rc = launched("pc_apply")) return rc;
  } else {
    launch_kernel(pc_apply_kernel<NMAX, 0>, grid, blk, (size_t)0, st, a, 0);
    if (int rc = launched("pc_apply")) return rc;
    launch_kernel(pc_finalize_kernel<NMAX>, dim3(1), dim3(32), (size_t)0, st, a, 1);
    if (int rc = launched("pc_finalize")) return rc;
    launch_kernel(pc_apply_kernel<NMAX, 1>, grid, blk, (size_t)0, st, a, 1);
    if (int rc = launched("pc_apply")) return rc;
  }
  if (a.corr_out) {
    launch_kernel(pc_finalize_kernel<NMAX>, dim3(1), dim3(32), (size_t)0, st, a, 2);
    if (int rc = launched("pc_finalize")) return rc;
  }
  return AE_OK;
}

inline int pc_ctas(int64_t n_el) {
  int64_t c = ceil_div64(n_el, kPcThreads);
  return (int)(c < 1 ? 1 : (c > kPcMaxCtas ? kPcMaxCtas : c));
}
inline int pc_nmax(int n) { return n <= 1 ? 1 : (n <= 4 ? 4 : (n <= 8 ? 8 : 16)); }
inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace
}  // namespace aedit

using namespace aedit;

extern "C" int64_t ae_pc_workspace_bytes(int n, int64_t n_el) {
  if (n < 1 || n > kPcMaxN || n_el < 1) return -1;
  const int nm = pc_nmax(n);
  return align256((int64_t)kPcMaxCtas * pc_nacc(nm) * 4) + align256((int64_t)n * n_el * 4) + align256(kS_total * 8);
}

extern "C" int ae_pc_perturb(const float* xt, int64_t xt_row_stride, const float* eigvecs, float amount, float sqrt_ab,
                             int mode, int n, float* x_batch, float* inp_out, int64_t n_el, ae_stream stream) {
  AE_CHECK_ARG(xt && x_batch && n >= 1 && n_el > 0, "ae_pc_perturb: null pointer / empty batch");
  AE_CHECK_ARG(mode >= 1 && mode <= 3, "ae_pc_perturb: mode must be 1 (both), 2 (text) or 3 (uncond)");
  PerturbArgs a{xt, xt_row_stride, eigvecs, amount, sqrt_ab, mode, n, x_batch, inp_out, n_el};
  dim3 grid((unsigned)pc_ctas(n_el), (unsigned)n);
  launch_kernel(pc_perturb_kernel, grid, dim3(256), (size_t)0, as_stream(stream), a);
  return launched("ae_pc_perturb");
}

extern "C" int ae_pc_subspace_step(const float* x0_pred, const float* x0_ref, const float* mask, const float* prev, int n,
                                   int64_t n_el, float cnst, float* eig_out, float* eig_scaled_out, float* norms_out,
                                   float* corr_out, void* workspace, int64_t workspace_bytes, ae_stream stream) {
  AE_CHECK_ARG(x0_pred && x0_ref && eig_out && norms_out && workspace, "ae_pc_subspace_step: null pointer");
  AE_CHECK_ARG(n >= 1 && n <= kPcMaxN, "ae_pc_subspace_step: n_ev=%d outside [1,%d]", n, kPcMaxN);
  AE_CHECK_ARG(n_el >= n, "ae_pc_subspace_step: fewer elements than directions");
  AE_CHECK_ARG(workspace_bytes >= ae_pc_workspace_bytes(n, n_el), "ae_pc_subspace_step: workspace too small");
  AE_CHECK_ARG(prev != eig_out && (corr_out == nullptr || prev != nullptr), "ae_pc_subspace_step: prev must not alias eig_out");
  const int nm = pc_nmax(n);
  PcArgs a{};
  a.x0_pred = x0_pred; a.x0_ref = x0_ref; a.mask = mask; a.prev = corr_out ? prev : nullptr;
  a.eig_out = eig_out; a.eig_scaled = eig_scaled_out; a.norms_out = norms_out; a.corr_out = corr_out;
  a.cnst = cnst; a.n = n; a.n_el = n_el; a.ctas = pc_ctas(n_el);
  char* p = static_cast<char*>(workspace);
  a.ws.partials = reinterpret_cast<float*>(p);
  p += align256((int64_t)kPcMaxCtas * pc_nacc(nm) * 4);
  a.ws.q1 = reinterpret_cast<float*>(p);
  p += align256((int64_t)n * n_el * 4);
  a.ws.small = reinterpret_cast<double*>(p);
  cudaStream_t st = as_stream(stream);
  switch (nm) {
    case 1: return pc_step_launch<1>(a, st);
    case 4: return pc_step_launch<4>(a, st);
    case 8: return pc_step_launch<8>(a, st);
    default: return pc_step_launch<16>(a, st);
  }
}

extern "C" int ae_pc_apply_drift(const float* xt_m1, const float* x0_pred, const float* latent, const float* shift_by,
                                 float std_dev_t, float sqrt_alpha_prev, float c_dir, float sqrt_ab_over_sqrt_1mab,
                                 int eta_positive, int use_shifted_x0_for_noisepred, int rows, int64_t n_el, float* out,
                                 ae_stream stream) {
  AE_CHECK_ARG(xt_m1 && x0_pred && shift_by && out && rows >= 1 && n_el > 0, "ae_pc_apply_drift: null pointer");
  AE_CHECK_ARG(!eta_positive || latent, "ae_pc_apply_drift: eta > 0 needs the variance noise");
  DriftArgs a{xt_m1, x0_pred, latent, shift_by, out, std_dev_t, sqrt_alpha_prev, c_dir, sqrt_ab_over_sqrt_1mab,
              eta_positive, use_shifted_x0_for_noisepred, n_el};
  dim3 grid((unsigned)pc_ctas(n_el), (unsigned)rows);
  launch_kernel(pc_apply_drift_kernel, grid, dim3(256), (size_t)0, as_stream(stream), a);
  return launched("ae_pc_apply_drift");
}
