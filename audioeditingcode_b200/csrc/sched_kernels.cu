// Scheduler-side elementwise kernels of the DDPM-inversion loop: K6/K7/K8 of SURVEY.md §2.2.
//   ae_sample_xts     models.py:67-83
//   ae_cfg_inv_step   inversion_utils.py:97-102 + models.py:85-117
//   ae_cfg_rev_step   inversion_utils.py:276-281 + models.py:119-158 + inversion_utils.py:308-315
//   ae_ddim_step      pc_drift.py:83-91 ([UPSTREAM] DDIMScheduler.step)
// Every arithmetic op is issued with an explicit round-to-nearest intrinsic (__fmul_rn, __fadd_rn, ...) in the
// order of the reference's eager PyTorch expression, so no FMA contraction happens and results are bit-identical
// to the reference's fp32 path (each PyTorch op rounds once).  Pure HBM-bound streaming: float4 where aligned.
#include <vector>
#include <unordered_map>

#include "common.cuh"

namespace aedit {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};
}  // namespace aedit

struct ae_sched {
  int N = 0;
  int pred_type = 0;
  int num_train = 0;
  std::vector<ae_sched_row> rows;
  std::unordered_map<long long, int> pos_of_t;
  ae_sched_row* d_rows = nullptr;
  float* d_eta = nullptr;  // [2*N]: c_dir[pos], sig[pos] supplied by the host (ae_sched_set_eta), or null
  bool has_eta = false;
};

using namespace aedit;

extern "C" const char* ae_last_error(void) { return g_err; }
extern "C" int ae_version(void) { return 200; }
extern "C" int ae_operand_dtype(void) { return AE_OPERAND_DTYPE; }
extern "C" int64_t ae_launch_count(void) { return g_launches.load(); }
extern "C" int ae_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

extern "C" int ae_sched_create(const float* ac, int T, float final_alpha, const int64_t* ts, int N, int pred_type,
                               ae_sched** out) {
  AE_CHECK_ARG(ac && ts && out && T > 0 && N > 0, "ae_sched_create: null/empty argument");
  AE_CHECK_ARG(pred_type == 0 || pred_type == 1, "ae_sched_create: pred_type must be 0 (epsilon) or 1 (v_prediction)");
  ae_sched* s = new ae_sched();
  s->N = N;
  s->pred_type = pred_type;
  s->num_train = T;
  s->rows.resize(N);
  const int step = T / N;  // models.py:96-97: num_train_timesteps // num_inference_steps (integer)
  for (int k = 0; k < N; ++k) {
    long long t = ts[k];
    if (t < 0 || t >= T) {
      delete s;
      return fail(AE_EINVAL, "ae_sched_create: timestep %lld out of range [0,%d)", t, T);
    }
    ae_sched_row r;
    r.t = (int32_t)t;
    r.prev_t = (int32_t)(t - step);
    // models.py:539-549 in the reference's fp32 operation order (volatile: forbid host FMA/extended precision)
    volatile float ab = ac[t];
    volatile float ap = r.prev_t >= 0 ? ac[r.prev_t] : final_alpha;
    volatile float beta_t = 1.0f - ab;
    volatile float beta_p = 1.0f - ap;
    volatile float ratio = beta_p / beta_t;
    volatile float q = ab / ap;
    volatile float one_m_q = 1.0f - q;
    volatile float var = ratio * one_m_q;
    r.alpha_bar_t = ab;
    r.alpha_prod_t_prev = ap;
    r.variance = var;
    r.sqrt_ab = sqrtf(ab);
    r.sqrt_1mab = sqrtf(beta_t);
    r.sqrt_ap = sqrtf(ap);
    r.sqrt_var = sqrtf(var);
    s->rows[k] = r;
    s->pos_of_t[t] = k;
  }
  cudaError_t e = cudaMalloc(&s->d_rows, sizeof(ae_sched_row) * N);
  if (e == cudaSuccess) e = cudaMemcpy(s->d_rows, s->rows.data(), sizeof(ae_sched_row) * N, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    // host-only table (no GPU): still usable for the index/scalar KATs
    cudaGetLastError();
    s->d_rows = nullptr;
  }
  *out = s;
  return AE_OK;
}
extern "C" int ae_sched_create_from_rows(const ae_sched_row* rows_h, int N, int pred_type, int num_train_timesteps,
                                         ae_sched** out) {
  AE_CHECK_ARG(rows_h && out && N > 0, "ae_sched_create_from_rows: null/empty argument");
  AE_CHECK_ARG(pred_type == 0 || pred_type == 1, "ae_sched_create_from_rows: bad pred_type");
  ae_sched* s = new ae_sched();
  s->N = N;
  s->pred_type = pred_type;
  s->num_train = num_train_timesteps;
  s->rows.assign(rows_h, rows_h + N);
  for (int k = 0; k < N; ++k) s->pos_of_t[rows_h[k].t] = k;
  cudaError_t e = cudaMalloc(&s->d_rows, sizeof(ae_sched_row) * N);
  if (e == cudaSuccess) e = cudaMemcpy(s->d_rows, s->rows.data(), sizeof(ae_sched_row) * N, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaGetLastError();
    s->d_rows = nullptr;
  }
  *out = s;
  return AE_OK;
}

extern "C" int ae_sched_set_eta(ae_sched* s, const float* c_dir_h, const float* sig_h) {
  AE_CHECK_ARG(s, "ae_sched_set_eta: null scheduler");
  if (!c_dir_h || !sig_h) {
    s->has_eta = false;
    return AE_OK;
  }
  if (!s->d_rows) return AE_OK;  // host-only table (no GPU)
  if (!s->d_eta) {
    cudaError_t e = cudaMalloc(&s->d_eta, sizeof(float) * 2 * s->N);
    if (e != cudaSuccess) return fail(AE_ECUDA, "ae_sched_set_eta: %s", cudaGetErrorString(e));
  }
  std::vector<float> tmp(2 * s->N);
  for (int k = 0; k < s->N; ++k) {
    tmp[k] = c_dir_h[k];
    tmp[s->N + k] = sig_h[k];
  }
  // synchronous copy: the table may be replaced between runs while kernels of the previous run are in flight
  cudaError_t e = cudaMemcpy(s->d_eta, tmp.data(), sizeof(float) * 2 * s->N, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(AE_ECUDA, "ae_sched_set_eta: %s", cudaGetErrorString(e));
  s->has_eta = true;
  return AE_OK;
}

extern "C" void ae_sched_destroy(ae_sched* s) {
  if (!s) return;
  if (s->d_rows) cudaFree(s->d_rows);
  if (s->d_eta) cudaFree(s->d_eta);
  delete s;
}
extern "C" int ae_sched_num_steps(const ae_sched* s) { return s ? s->N : 0; }
extern "C" int ae_sched_row_h(const ae_sched* s, int pos, ae_sched_row* out) {
  AE_CHECK_ARG(s && out && pos >= 0 && pos < s->N, "ae_sched_row_h: bad position %d", pos);
  *out = s->rows[pos];
  return AE_OK;
}
extern "C" int ae_sched_pos_of_t(const ae_sched* s, int64_t t) {
  if (!s) return -1;
  auto it = s->pos_of_t.find(t);
  return it == s->pos_of_t.end() ? -1 : it->second;
}

// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxP = 8;

// mu_xt of models.py:91-109 / :132-150 for one element (pred_type 0 = epsilon, 1 = v_prediction)
__device__ __forceinline__ float ddpm_mu(const ae_sched_row& r, int pred_type, float c_dir, float xt, float eps) {
  float x0p, e;
  if (pred_type == 0) {
    x0p = __fdiv_rn(__fsub_rn(xt, __fmul_rn(r.sqrt_1mab, eps)), r.sqrt_ab);
    e = eps;
  } else {
    x0p = __fsub_rn(__fmul_rn(r.sqrt_ab, xt), __fmul_rn(r.sqrt_1mab, eps));
    e = __fadd_rn(__fmul_rn(r.sqrt_ab, eps), __fmul_rn(r.sqrt_1mab, xt));
  }
  float dir = __fmul_rn(c_dir, e);
  return __fadd_rn(__fmul_rn(r.sqrt_ap, x0p), dir);
}

// (1 - alpha_prod_t_prev - eta*variance) ** 0.5   (note eta, not eta**2: models.py:107,148)
__device__ __forceinline__ float dir_coeff(const ae_sched_row& r, float eta) {
  return __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, r.alpha_prod_t_prev), __fmul_rn(eta, r.variance)));
}

__global__ void sample_xts_kernel(const ae_sched_row* __restrict__ rows, int N, const float* __restrict__ x0,
                                  const float* __restrict__ noise, float* __restrict__ xts, int64_t n_el) {
  pdl_trigger();
  pdl_wait();
  // grid.y = slot 0..N (slot 0 copies x0); slot s>0 <-> pos = N - s, draw k = N-1-pos = s-1
  const int s = blockIdx.y;
  float a = 1.0f, b = 0.0f;
  const float* nz = nullptr;
  if (s > 0) {
    const ae_sched_row r = rows[N - s];
    a = r.sqrt_ab;
    b = r.sqrt_1mab;
    nz = noise + (int64_t)(s - 1) * n_el;
  }
  float* dst = xts + (int64_t)s * n_el;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x0[i];
    if (s > 0) v = __fadd_rn(__fmul_rn(v, a), __fmul_rn(nz[i], b));
    dst[i] = v;
  }
}

struct InvArgs {
  const float* eta_tab;  // [2N] c_dir | sig, or null (computed in-kernel with IEEE ops)
  const ae_sched_row* rows;
  int N, pos0, pred_type, P, numerical_fix;
  float eta;
  const float* eps_u;
  int64_t ld_eps_u;
  const float* eps_c;
  int64_t ld_eps_c;
  const float* cfg_map;
  const float* xt_src;
  float* xts;
  float* zs;
  int64_t n_el;
};

__global__ void cfg_inv_step_kernel(InvArgs a) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.y;
  const int pos = a.pos0 + j;
  const int idx = a.N - pos - 1;  // inversion_utils.py:75
  const ae_sched_row r = a.rows[pos];
  const float c_dir = a.eta_tab ? a.eta_tab[pos] : dir_coeff(r, a.eta);
  const float sig = a.eta_tab ? a.eta_tab[a.N + pos] : __fmul_rn(a.eta, r.sqrt_var);
  const float* eu = a.eps_u + (int64_t)j * a.ld_eps_u;
  const float* xt_p = a.xt_src + (int64_t)(idx + 1) * a.n_el;
  float* xtm1_p = a.xts + (int64_t)idx * a.n_el;
  float* z_p = a.zs + (int64_t)idx * a.n_el;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_el; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = eu[i];
    float eps = u;
    if (a.P > 0) {
      float acc = 0.0f;
      for (int p = 0; p < a.P; ++p) {
        const float c = a.eps_c[((int64_t)j * a.P + p) * a.ld_eps_c + i];
        const float term = __fmul_rn(a.cfg_map[(int64_t)p * a.n_el + i], __fsub_rn(c, u));
        acc = (p == 0) ? term : __fadd_rn(acc, term);
      }
      eps = __fadd_rn(u, acc);
    }
    const float xt = xt_p[i];
    const float mu = ddpm_mu(r, a.pred_type, c_dir, xt, eps);
    const float z = __fdiv_rn(__fsub_rn(xtm1_p[i], mu), sig);
    z_p[i] = z;
    if (a.numerical_fix) xtm1_p[i] = __fadd_rn(mu, __fmul_rn(sig, z));
  }
}

struct RevArgs {
  const float* eta_tab;
  int N;
  const ae_sched_row* rows;
  int pos;
  const int32_t* d_pos;
  int pred_type, P;
  float eta;
  const float* eps_u;
  const float* eps_c;
  const float* cfg_map;
  const float* xt;
  const float* z;
  float* xt_out;
  const float* masks;
  const float* xT_fix;
  float fix_alpha[kMaxP];
  int do_fix;
  int64_t n_el;
};

__global__ void cfg_rev_step_kernel(RevArgs a) {
  pdl_trigger();
  pdl_wait();
  const int pos = a.d_pos ? *a.d_pos : a.pos;
  const ae_sched_row r = a.rows[pos];
  const float c_dir = a.eta_tab ? a.eta_tab[pos] : dir_coeff(r, a.eta);
  const float sig = a.eta_tab ? a.eta_tab[a.N + pos] : __fmul_rn(a.eta, r.sqrt_var);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_el; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = a.eps_u[i];
    float eps = u;
    if (a.P > 0) {
      float acc = 0.0f;
      for (int p = 0; p < a.P; ++p) {
        const float term = __fmul_rn(a.cfg_map[(int64_t)p * a.n_el + i], __fsub_rn(a.eps_c[(int64_t)p * a.n_el + i], u));
        acc = (p == 0) ? term : __fadd_rn(acc, term);
      }
      eps = __fadd_rn(u, acc);
    }
    float prev = ddpm_mu(r, a.pred_type, c_dir, a.xt[i], eps);
    if (a.eta > 0.0f) prev = __fadd_rn(prev, __fmul_rn(sig, a.z[i]));
    if (a.do_fix) {
      // inversion_utils.py:311-315: sum_p masks[p] * (xt*(1-a_p) + a_p*xT_fix)
      const float xf = a.xT_fix[i];
      float acc = 0.0f;
      for (int p = 0; p < a.P; ++p) {
        const float ap = a.fix_alpha[p];
        const float blended = __fadd_rn(__fmul_rn(prev, __fsub_rn(1.0f, ap)), __fmul_rn(ap, xf));
        const float term = __fmul_rn(a.masks[(int64_t)p * a.n_el + i], blended);
        acc = (p == 0) ? term : __fadd_rn(acc, term);
      }
      prev = acc;
    }
    a.xt_out[i] = prev;
  }
}

__global__ void ddim_step_kernel(const ae_sched_row* __restrict__ rows, int pos, int pred_type, float eta, float cfg,
                                 const float* __restrict__ eps_u, const float* __restrict__ eps_c,
                                 const float* __restrict__ sample, const float* __restrict__ vnoise,
                                 float* __restrict__ prev_out, float* __restrict__ x0_out, int64_t n_el) {
  pdl_trigger();
  pdl_wait();
  const ae_sched_row r = rows[pos];
  const float std_t = __fmul_rn(eta, r.sqrt_var);
  const float c_dir = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, r.alpha_prod_t_prev), __fmul_rn(std_t, std_t)));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = eps_u[i];
    float mo = u;
    if (eps_c) mo = __fadd_rn(u, __fmul_rn(cfg, __fsub_rn(eps_c[i], u)));  // pc_drift.py:83
    const float x = sample[i];
    float x0p, e;
    if (pred_type == 0) {
      x0p = __fdiv_rn(__fsub_rn(x, __fmul_rn(r.sqrt_1mab, mo)), r.sqrt_ab);
      e = mo;
    } else {
      x0p = __fsub_rn(__fmul_rn(r.sqrt_ab, x), __fmul_rn(r.sqrt_1mab, mo));
      e = __fadd_rn(__fmul_rn(r.sqrt_ab, mo), __fmul_rn(r.sqrt_1mab, x));
    }
    float prev = __fadd_rn(__fmul_rn(r.sqrt_ap, x0p), __fmul_rn(c_dir, e));
    if (eta > 0.0f) prev = __fadd_rn(prev, __fmul_rn(std_t, vnoise[i]));
    prev_out[i] = prev;
    if (x0_out) x0_out[i] = x0p;
  }
}

inline int grid_for(int64_t n, int threads, int rows) {
  int64_t blocks = ceil_div64(n, threads);
  int64_t cap = 148LL * 8 / (rows > 0 ? 1 : 1);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int ae_sample_xts(const ae_sched* s, const float* x0, const float* noise, float* xts, int64_t n_el,
                             ae_stream stream) {
  AE_CHECK_ARG(s && s->d_rows && x0 && noise && xts && n_el > 0, "ae_sample_xts: bad argument");
  dim3 grid(grid_for(n_el, 256, s->N + 1), s->N + 1);
  launch_kernel(sample_xts_kernel, dim3(grid), dim3(256), (size_t)(0), as_stream(stream), s->d_rows, s->N, x0, noise, xts, n_el);
  return launched("ae_sample_xts");
}

extern "C" int ae_cfg_inv_step(const ae_sched* s, int pos0, int count, float eta, const float* eps_u, int64_t ld_eps_u,
                               const float* eps_c, int64_t ld_eps_c, int P, const float* cfg_map, const float* xt_src,
                               float* xts, float* zs, int numerical_fix, int64_t n_el, ae_stream stream) {
  AE_CHECK_ARG(s && s->d_rows, "ae_cfg_inv_step: scheduler has no device table");
  AE_CHECK_ARG(pos0 >= 0 && count > 0 && pos0 + count <= s->N, "ae_cfg_inv_step: positions [%d,%d) outside [0,%d)", pos0,
               pos0 + count, s->N);
  AE_CHECK_ARG(eps_u && xt_src && xts && zs && n_el > 0, "ae_cfg_inv_step: null pointer");
  AE_CHECK_ARG(P >= 0 && P <= kMaxP && (P == 0 || (eps_c && cfg_map)), "ae_cfg_inv_step: bad P=%d", P);
  AE_CHECK_ARG(eta > 0.0f, "ae_cfg_inv_step: eta must be > 0 (z is divided by eta*sqrt(var))");
  InvArgs a{s->has_eta ? s->d_eta : nullptr, s->d_rows, s->N, pos0, s->pred_type, P, numerical_fix, eta, eps_u, ld_eps_u, eps_c, ld_eps_c,
            cfg_map, xt_src, xts, zs, n_el};
  dim3 grid(grid_for(n_el, 256, count), count);
  launch_kernel(cfg_inv_step_kernel, dim3(grid), dim3(256), (size_t)(0), as_stream(stream), a);
  return launched("ae_cfg_inv_step");
}

extern "C" int ae_cfg_rev_step(const ae_sched* s, int pos, const int32_t* d_pos, float eta, const float* eps_u,
                               const float* eps_c, int P, const float* cfg_map, const float* xt, const float* z,
                               float* xt_out, const float* masks, const float* fix_alpha_h, const float* xT_fix,
                               int64_t n_el, ae_stream stream) {
  AE_CHECK_ARG(s && s->d_rows, "ae_cfg_rev_step: scheduler has no device table");
  AE_CHECK_ARG(d_pos || (pos >= 0 && pos < s->N), "ae_cfg_rev_step: bad position %d", pos);
  AE_CHECK_ARG(eps_u && xt && xt_out && n_el > 0, "ae_cfg_rev_step: null pointer");
  AE_CHECK_ARG(P >= 0 && P <= kMaxP && (P == 0 || (eps_c && cfg_map)), "ae_cfg_rev_step: bad P=%d", P);
  AE_CHECK_ARG(eta == 0.0f || z, "ae_cfg_rev_step: eta > 0 needs the noise map z");
  RevArgs a;
  a.eta_tab = s->has_eta ? s->d_eta : nullptr;
  a.N = s->N;
  a.rows = s->d_rows;
  a.pos = pos;
  a.d_pos = d_pos;
  a.pred_type = s->pred_type;
  a.P = P;
  a.eta = eta;
  a.eps_u = eps_u;
  a.eps_c = eps_c;
  a.cfg_map = cfg_map;
  a.xt = xt;
  a.z = z;
  a.xt_out = xt_out;
  a.masks = masks;
  a.xT_fix = xT_fix;
  a.do_fix = 0;
  for (int p = 0; p < kMaxP; ++p) a.fix_alpha[p] = 0.0f;
  if (fix_alpha_h) {
    AE_CHECK_ARG(masks && xT_fix && P > 0, "ae_cfg_rev_step: mask fix needs masks, xT_fix and P > 0");
    for (int p = 0; p < P; ++p) a.fix_alpha[p] = fix_alpha_h[p];
    a.do_fix = 1;
  }
  a.n_el = n_el;
  launch_kernel(cfg_rev_step_kernel, dim3(grid_for(n_el, 256, 1)), dim3(256), (size_t)(0), as_stream(stream), a);
  return launched("ae_cfg_rev_step");
}

extern "C" int ae_ddim_step(const ae_sched* s, int pos, float eta, float cfg_scale, const float* eps_u,
                            const float* eps_c, const float* sample, const float* vnoise, float* prev_sample,
                            float* pred_x0, int64_t n_el, ae_stream stream) {
  AE_CHECK_ARG(s && s->d_rows && pos >= 0 && pos < s->N, "ae_ddim_step: bad scheduler/position");
  AE_CHECK_ARG(eps_u && sample && prev_sample && n_el > 0, "ae_ddim_step: null pointer");
  AE_CHECK_ARG(eta == 0.0f || vnoise, "ae_ddim_step: eta > 0 needs variance_noise");
  launch_kernel(ddim_step_kernel, dim3(grid_for(n_el, 256, 1)), dim3(256), (size_t)(0), as_stream(stream), s->d_rows, pos, s->pred_type, eta, cfg_scale,
                                                                          eps_u, eps_c, sample, vnoise, prev_sample,
                                                                          pred_x0, n_el);
  return launched("ae_ddim_step");
}
