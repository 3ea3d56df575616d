/* libaedit.so — C ABI of the B200-native DDPM-inversion / CFG-denoising hot path.
 *
 * The reference (HilaManor/AudioEditingCode) has no FFI: its "plugin" boundary is the Python duck-typed
 * `PipelineWrapper` protocol of code/models.py:14-158 that the loops in
 * code/ddm_inversion/inversion_utils.py and code/pc_drift.py call.  The Python package
 * `audioeditingcode_b200` mirrors that protocol; every piece of device math it performs goes through the
 * entry points below (plain pointers and sizes, no torch types).  Each entry point cites the reference
 * lines whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless suffixed _h (host)
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises, never
 *     allocates device memory (the only allocations are in *_create), and is CUDA-graph capturable
 *   - return 0 on success, negative AE_E* otherwise; ae_last_error() returns a thread-local message
 *   - activations are channels-last: images [B,H,W,C], token sequences [B,T,C]; "f32" = float.  Latents / noise maps of
 *     the scheduler kernels are plain contiguous fp32 of any layout.
 *   - "bf16" in parameter names below means THE LIBRARY'S 16-BIT TENSOR-CORE OPERAND TYPE, reported by
 *     ae_operand_dtype(): 1 = IEEE fp16 (libaedit.so, the default build: same tcgen05 rate, 8x smaller operand rounding
 *     error, conversions saturate at +-65504) or 0 = bfloat16 (libaedit_bf16.so, built with -DAE_OPERAND_BF16).
 *     Accumulation, residual streams, norm statistics and all scheduler state are fp32 in both builds.
 */
#ifndef AEDIT_H_
#define AEDIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AE_OK 0
#define AE_EINVAL (-1)   /* bad argument / unsupported shape */
#define AE_ECUDA (-2)    /* CUDA runtime / driver error */
#define AE_EUNSUPPORTED (-3)

typedef void* ae_stream; /* cudaStream_t */

const char* ae_last_error(void);
int ae_version(void);
/* 16-bit operand type of this build: 0 bfloat16, 1 IEEE fp16 (see conventions above) */
int ae_operand_dtype(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t ae_launch_count(void);
/* 1 if the current device is compute capability 10.x */
int ae_device_ok(void);
/* ---- Settings.  Every ae_set_* below changes a setting OF THE CALLING HOST THREAD only (thread_local state): it
 * affects the kernels that thread launches (or captures into a CUDA graph) afterwards, on whatever stream.  The host
 * toggles some of them around graph captures (launch priority, PDL families, shared-SM rings); threads driving other
 * streams are not affected, and a thread that never calls a setter runs with the documented defaults.  Work on several
 * streams from ONE thread is unaffected by construction — a setting is read on the host when a kernel is launched. */
/* Programmatic dependent launch: a kernel launched with it may start (prologue, weight prefetch) before its
 * predecessor in the stream has drained, and waits for it before touching dependent global memory. */
void ae_set_pdl(int mode); /* 0 off, 1 every kernel, 2 GEMM kernels only (default) */

/* Launch priority (cudaLaunchAttributePriority) attached to every kernel this library launches while it is set; 0 (the
 * default) leaves the stream's own priority.  The host captures the reverse-process U-Net graph (reference
 * inversion_utils.py:229-320, a strictly sequential chain of sub-wave kernels) with the device's highest priority,
 * so that it can run CONCURRENTLY with the throughput-bound forward-process chunks (inversion_utils.py:69-131) of
 * the same clip and still get the next free SM slots. */
void ae_set_launch_priority(int prio);
/* PDL mode 2 extension: kernel families that are also launched with the PDL attribute (1 GroupNorm statistics,
 * 2 GroupNorm apply, 4 LayerNorm, 8 attention).  Default 0. */
void ae_set_pdl_extra(int mask);
/* GroupNorm tensors of at least this many bytes (default 8 MiB) take the streaming apply kernel (same bits). */
void ae_set_gn_stream_min_bytes(int64_t bytes);
/* Multi-wave 128-wide GEMM grids whose K loop has at most this many 64-wide blocks use a 2-stage operand ring (three
 * CTAs per SM instead of two).  0 = never.  Same bits. */
void ae_set_shallow_kblocks(int kb);
/* GEMM grids of at least this many 128 x BN output tiles (no K split, no batch) run as persistent CTAs, one per SM,
 * with the accumulator double-buffered in TMEM so that a tile's epilogue overlaps the next tile's main loop — for the
 * shapes where that measured faster (linears with a bf16 / GEGLU output or >= 9 K blocks).  Default 296; 0 = never.
 * Same bits as the one-CTA-per-tile kernel. */
void ae_set_persistent_min_tiles(int tiles);
/* Tile model constants: what a split-K reduce pass is charged (launch nanoseconds, bytes per microsecond). */
void ae_set_tile_model_reduce(int launch_ns, int bytes_per_us);
/* DIAGNOSTIC ONLY: drop the launches of kernel families (1 GEMM, 2 split-K reduce, 4 GroupNorm statistics,
 * 8 GroupNorm apply, 16 LayerNorm, 32 attention) to measure a family's marginal cost inside a captured graph
 * (tools/kernel_share.py).  Outputs are meaningless while the mask is non-zero. */
void ae_set_skip_mask(int mask);
/* 1: the GEMMs launched (captured) from now on will share the SMs with a throughput-bound grid of another stream;
 * sub-wave grids then take the deepest operand ring that fits beside one resident ~100 KB CTA instead of the 6-stage
 * ring.  0 (default): sub-wave grids own their SM.  Results are bit-identical in both modes. */
void ae_set_shared_sm(int on);
/* the current device's highest stream / launch priority (numerically lowest value; 0 when there is one level) */
int ae_greatest_priority(void);

/* CTA budget the automatic split-K heuristic fills (default 148 = one per SM).  Callers that run two dependency chains
 * concurrently on forked streams lower it so that the chains share the SMs instead of queueing behind each other. */
void ae_set_splitk_ctas(int ctas);

/* ae_gemm epilogue: 1 (default) = coalesced, shared-memory-staged epilogue where it measured faster (residual
 * read, or fp32 tile written by a multi-wave grid) and the operands are 16-byte tileable; 2 = whenever legal;
 * 0 = always the row-per-thread epilogue.  All three give identical bits (tests/test_gpu_kernels.py). */
void ae_set_fast_epilogue(int mode);

/* 1 (default): sub-wave GEMM grids choose tile width and K split jointly from an operand-bytes-per-SM model;
 * 0: fixed rule (128-wide tiles, split only for >= 12 K blocks) — kept for A/B measurements. */
void ae_set_tile_model(int on);



/* ------------------------------------------------------------------------------------------------
 * Scheduler table  (code/models.py:85-158, :539-549; integer index math of
 * code/ddm_inversion/inversion_utils.py:75,222-224 and models.py:76-80,96-97,123-124)
 *
 * Built on the host from alphas_cumprod [T] (fp32, exactly the scheduler's tensor), final_alpha_cumprod and
 * the descending `timesteps` [N].  Row k (k = position in `timesteps`) holds, in the reference's fp32
 * operation order:  t, prev_t, alpha_bar_t, alpha_prod_t_prev, variance,
 *   sqrt_ab = ab**0.5, sqrt_1mab = (1-ab)**0.5, sqrt_ap = ap**0.5, sqrt_var = var**0.5
 * `eta` dependent terms are formed at call time (eta is per call in the reference).
 * pred_type: 0 = epsilon, 1 = v_prediction.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ae_sched ae_sched;

typedef struct {
  int32_t t, prev_t;
  float alpha_bar_t, alpha_prod_t_prev, variance;
  float sqrt_ab, sqrt_1mab, sqrt_ap, sqrt_var;
} ae_sched_row;

int ae_sched_create(const float* alphas_cumprod_h, int T, float final_alpha_cumprod, const int64_t* timesteps_h, int N,
                    int pred_type, ae_sched** out);
/* Same table from rows computed by the caller.  The Python host computes the scalars with the SAME torch CPU ops the
 * reference uses (torch's CPU `x ** 0.5` is not always the IEEE-rounded sqrtf ae_sched_create uses: 1 scalar in
 * ~500 differs by an ulp), so that results stay bit-identical to the reference's. */
int ae_sched_create_from_rows(const ae_sched_row* rows_h, int N, int pred_type, int num_train_timesteps, ae_sched** out);
/* Optional per-position eta terms from the host (same motivation): c_dir[pos] = (1 - alpha_prod_t_prev - eta*var)**0.5
 * and sig[pos] = eta*var**0.5 (models.py:107,111,155).  NULL pointers revert to in-kernel IEEE evaluation from `eta`. */
int ae_sched_set_eta(ae_sched*, const float* c_dir_h, const float* sig_h);
void ae_sched_destroy(ae_sched*);
int ae_sched_num_steps(const ae_sched*);
/* host copy of row `pos` (pos indexes `timesteps`) */
int ae_sched_row_h(const ae_sched*, int pos, ae_sched_row* out);
/* position of timestep value t in `timesteps`, or -1   (t_to_idx of inversion_utils.py:68) */
int ae_sched_pos_of_t(const ae_sched*, int64_t t);

/* models.py:67-83 sample_xts_from_x0:  xts[0] = x0;  xts[N - pos] = x0*sqrt(ab[t_pos]) + noise[k]*sqrt(1-ab[t_pos])
 * where k = N-1-pos is the draw order of the reference (ascending t).  noise: [N, n_el], xts: [N+1, n_el]. */
int ae_sample_xts(const ae_sched*, const float* x0, const float* noise, float* xts, int64_t n_el, ae_stream stream);

/* Fused CFG combine (inversion_utils.py:97-102) + get_zs_from_xts (models.py:85-117) for `count` consecutive
 * loop positions pos0 .. pos0+count-1 (count > 1 = the timestep-batched forward process).  For position pos,
 * idx = N - pos - 1 (inversion_utils.py:75):
 *      xt    = xts_in[idx+1]           (read from `xt_src`, usually == xts)
 *      eps   = eps_u[j] + sum_p cfg_map[p] * (eps_c[j,p] - eps_u[j])      j = pos - pos0;  P == 0: eps = eps_u[j]
 *      z     = (xts[idx] - mu) / (eta*sqrt(var));   zs[idx] = z;
 *      numerical_fix: xts[idx] = mu + eta*sqrt(var)*z
 * eps_u: [count, n_el] with row stride ld_eps_u; eps_c: [count, P, n_el] (row stride ld_eps_c per (j,p) row);
 * cfg_map: [P, n_el].  xt_src and xts may alias (sequential semantics are then the caller's business). */
int ae_cfg_inv_step(const ae_sched*, int pos0, int count, float eta, const float* eps_u, int64_t ld_eps_u,
                    const float* eps_c, int64_t ld_eps_c, int P, const float* cfg_map, const float* xt_src, float* xts,
                    float* zs, int numerical_fix, int64_t n_el, ae_stream stream);

/* Fused CFG combine (inversion_utils.py:276-281) + reverse_step_with_custom_noise (models.py:119-158) +
 * optional multi-prompt mask "fix" (inversion_utils.py:308-315) for ONE loop position `pos`.
 * If d_pos != NULL the position is read from device memory instead (graph replay).
 *      xt_out = mu(xt, eps) + eta*sqrt(var)*z            (eta > 0; z may be NULL when eta == 0)
 * fix (n_fix_p > 0): xt_out = sum_p mask[p] * (xt_out*(1-a_p) + a_p*xT_fix), a_p = fix_alpha[p] (0 = no fix for p) */
int ae_cfg_rev_step(const ae_sched*, int pos, const int32_t* d_pos, float eta, const float* eps_u, const float* eps_c,
                    int P, const float* cfg_map, const float* xt, const float* z, float* xt_out, const float* masks,
                    const float* fix_alpha_h, const float* xT_fix, int64_t n_el, ae_stream stream);

/* [UPSTREAM] DDIMScheduler.step as used by code/pc_drift.py:89 (std = eta*sqrt(var), direction uses std**2):
 * writes prev_sample and pred_original_sample for `B` rows sharing timestep position pos. cfg: scalar guidance. */
int ae_ddim_step(const ae_sched*, int pos, float eta, float cfg_scale, const float* eps_u, const float* eps_c,
                 const float* sample, const float* variance_noise, float* prev_sample, float* pred_x0, int64_t n_el,
                 ae_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Unsupervised principal-direction editing (code/pc_drift.py), all fp32, `n` direction rows of `n_el` elements.
 */
/* pc_drift.py:41-42, :64-80 — the input of forward_directional written as the 2n-row CFG batch of one U-Net launch:
 *   inp[i]         = xt[i] + (amount * eigvecs[i]) * sqrt_ab          (eigvecs == NULL: inp = xt)
 *   x_batch[i]     = mode in {1 both, 3 uncond} ? inp[i] : xt[i]      rows evaluated with the unconditional embedding
 *   x_batch[n + i] = mode in {1 both, 2 text}   ? inp[i] : xt[i]      rows evaluated with the text embedding
 * xt rows are xt_row_stride elements apart (0 = one row shared by all directions); inp_out (optional) [n, n_el] is the
 * `sample` later handed to ae_ddim_step. */
int ae_pc_perturb(const float* xt, int64_t xt_row_stride, const float* eigvecs, float amount, float sqrt_ab, int mode,
                  int n, float* x_batch, float* inp_out, int64_t n_el, ae_stream stream);

/* pc_drift.py:148-193 — one subspace-iteration update after the U-Net evaluation of the n perturbed rows:
 *   Ab_i = x0_pred[i]*mask - x0_ref;  norms_out[i] = ||Ab_i[mask != 0]||;  V_i = Ab_i / norm_i * mask
 *   n > 1: Q = torch.linalg.qr(V^T).Q — computed as CholeskyQR2 (Gram pass / n x n factorisation / transform pass, twice)
 *          with LAPACK's Householder column-sign convention, `Q *= -1 if prod(diag R) < 0` (:164-166) and the rows then
 *          re-ordered by norms_out descending, stable (:172-174);  n == 1: eig = V_0 (:175-177)
 *   corr_out[k] = <prev[k], eig_out[k]> (:181-182; needs prev, which must not alias eig_out)
 *   eig_scaled_out = cnst * eig_out (:193, the next iteration's perturbation)
 * mask [n_el] (NULL = ones) and x0_ref [n_el] are shared by the n rows.  Deterministic: per-CTA partial sums are combined
 * in a fixed order.  workspace: ae_pc_workspace_bytes(n, n_el) bytes, no initialisation needed.  1 <= n <= 16. */
int64_t ae_pc_workspace_bytes(int n, int64_t n_el);
int ae_pc_subspace_step(const float* x0_pred, const float* x0_ref, const float* mask, const float* prev, int n,
                        int64_t n_el, float cnst, float* eig_out, float* eig_scaled_out, float* norms_out,
                        float* corr_out, void* workspace, int64_t workspace_bytes, ae_stream stream);

/* pc_drift.py:232-278 — apply_drift for `rows` samples sharing the shift `shift_by` [n_el] = sum_k amount*sqrt(eigval_k)*
 * eigvec_k; scalars are the host-evaluated std_dev_t = eta*sqrt(var), sqrt(alpha_prod_t_prev),
 * sqrt(1 - alpha_prod_t_prev - std_dev_t^2) and sqrt(alpha_prod_t)/sqrt(1 - alpha_prod_t). */
int ae_pc_apply_drift(const float* xt_m1, const float* x0_pred, const float* latent, const float* shift_by,
                      float std_dev_t, float sqrt_alpha_prev, float c_dir, float sqrt_ab_over_sqrt_1mab, int eta_positive,
                      int use_shifted_x0_for_noisepred, int rows, int64_t n_el, float* out, ae_stream stream);

/* ------------------------------------------------------------------------------------------------
 * U-Net building blocks (the math of `self.model.unet.*` called from models.py:231-388 / :772-894; in-tree
 * statement: code/audioldm/latent_diffusion/openaimodel.py, attention.py, util.py)
 * ------------------------------------------------------------------------------------------------ */

/* D[M,N] = alpha * A[M,K] · W[N,K]^T  (+bias[n]) (+rowbias[m / rows_per_group, n]) (+residual[m,n])  -> act -> out
 * A, W: bf16, K contiguous.  tcgen05 tensor-core tiles (128 x BN x 64) fed by TMA, fp32 accumulation in TMEM.
 * conv != 0: A is a channels-last image [B,H,W,C] and the K dimension is gathered on the fly by TMA
 *            (implicit GEMM, zero padding by out-of-bounds fill): K = kh*kw*C ordered (kh,kw,c), stride 1,
 *            output positions = input positions ("same" size: pad = dil*(k-1)/2).  M = B*H*W.
 * batch > 1: independent problems z = 0..batch-1 with element strides (A, W, outputs, residual). */
typedef struct {
  const void* A;
  int64_t lda;
  const void* W;
  int64_t ldw;
  int32_t M, N, K;
  int32_t batch;
  int64_t strideA, strideW, stride_out, stride_res;
  const float* bias;
  const float* rowbias;
  int64_t ld_rowbias;
  int32_t rows_per_group;
  const float* residual;
  int64_t ld_res;
  float* out_f32;
  int64_t ld_out_f32;
  void* out_bf16;
  int64_t ld_out_bf16;
  int32_t act; /* 3 = grouped softmax (see sm_* below); 0 none, 1 SiLU, 2 GEGLU: W rows interleaved in blocks of 32 (16 value rows, 16 gate rows); the
                  output has N/2 columns: out[16q+i] = acc[32q+i] * gelu(acc[32q+16+i])   (attention.py:37-44) */
  float alpha;
  /* implicit convolution */
  int32_t conv;
  int32_t B, H, W_, C, kh, kw, dil_h, dil_w;
  int32_t force_bn; /* 0 = auto tile width, else 32/64/128 */
  /* split-K (small-M layers): fp32 partial tiles go to this workspace, a second kernel reduces them in fixed order
   * and applies the epilogue.  NULL = never split.  force_split: 0 auto, 1 never, >1 exactly that many K slices. */
  float* splitk_ws;
  int64_t splitk_ws_bytes;
  int32_t force_split;
  int32_t w_dynamic;    /* 1 if W is written by a preceding kernel on the stream (e.g. K or V^T of an unfused attention):
                           disables the early W prefetch that otherwise overlaps the previous kernel's tail */
  int32_t force_stages; /* 0 auto (deep 6-stage ring for grids <= 160 CTAs, else 3 stages x 2-3 CTAs/SM), 3 or 6 */
  /* GroupNorm statistics of the OUTPUT, produced by the epilogue (or the split-K reduce) instead of a separate pass
   * over the tensor (the GroupNorm that follows — openaimodel.py:213-216,238-239 — needs per-sample sums over
   * positions): colstats[(sample*N + n)*2 + {0,1}] += fixed-point (sum * 2^28, sum of squares * 2^24) of column n over
   * the rows of the sample, sample = row / cs_rows_per_sample.  int64 accumulators, ZEROED by the caller; integer
   * atomics make the totals independent of CTA arrival order.  Needs batch == 1, act != 2, fp32 output, N % 4 == 0,
   * cs_rows_per_sample % 32 == 0.  Consumed by ae_groupnorm_cs.  NULL = off. */
  int64_t* colstats;
  int32_t cs_rows_per_sample;
  int32_t force_persistent; /* 0 auto (ae_set_persistent_min_tiles), 1 persistent kernel, -1 one CTA per tile */
  /* act == 3 — grouped softmax epilogue.  Cross-attention against FROZEN text (attention.py:234-262 with context =
   * the prompt embedding, constant over all denoising steps) folds into two small GEMMs:
   *   scores[m, (r,h,l)] = LN(x)[m,:] . KW[(r,h,l),:],  KW[(r,h,l), c] = scale * sum_j K_r[l, h*d+j] * Wq[h*d+j, c]
   *   out[m, c]          = sum_(r,h,l) P[m,(r,h,l)] * VW[c,(r,h,l)] (+ bias + residual),  VW = Wo_h . V_r,h^T
   * where P = softmax over l within each (r, h) group for r = sm_slot[m / sm_rows] (the sample's text row) and 0 for the
   * other text rows.  This call computes `scores` and writes P (bf16): sm_L = padded keys per group (8, 16 or 32),
   * sm_block = heads * sm_L columns per text row, sm_bias (optional) additive fp32 [rows, sm_L] (key mask: 0 keep,
   * -10000 discard as models.py:204-210, -1e30 for padding keys). */
  int32_t sm_L, sm_block, sm_rows;
  const int32_t* sm_slot;
  const float* sm_bias;
} ae_gemm_args;
int ae_gemm(const ae_gemm_args*, ae_stream stream);
/* 1 if the implicit-conv fast path supports this geometry (else use ae_im2col + plain GEMM) */
int ae_gemm_conv_supported(int B, int H, int W, int C);

/* Explicit patch gather for the convolutions the TMA path does not cover (stride 2, asymmetric padding, C%64!=0):
 * out[m, (i*kw + j)*C + c] = in[b, ho*stride - pad_t + i*dil, wo*stride - pad_l + j*dil, c]  (0 outside), bf16 out.
 * in: f32 or bf16 [B,H,W,C];  out: [B*Ho*Wo, ld_out] with ld_out >= kh*kw*C (tail columns zeroed). */
int ae_im2col(const void* in, int in_is_bf16, int B, int H, int W, int C, int kh, int kw, int stride, int dil,
              int pad_t, int pad_l, int Ho, int Wo, void* out_bf16, int64_t ld_out, ae_stream stream);

/* GroupNorm (+ optional SiLU) over channels-last fp32 input, bf16 output (openaimodel.py:213-216,238-239; util.py:240).
 * The input may be a virtual channel concatenation of two tensors (up-block skip concat, openaimodel.py:845):
 * x = cat([x1 (C1 ch), x2 (C2 ch)], channel).  HW = spatial positions per sample.  raw_out (optional): bf16 copy of x.
 * workspace: ae_groupnorm_workspace_bytes(B, groups) bytes, ZERO-initialised once by the caller (it holds the
 * per-sample arrival counters, which the kernel re-arms itself). */
int64_t ae_groupnorm_workspace_bytes(int B, int groups);
int ae_groupnorm(const float* x1, int C1, const float* x2, int C2, int B, int64_t HW, int groups, float eps,
                 const float* gamma, const float* beta, int silu, void* out_bf16, void* raw_out_bf16,
                 float* cat_out_f32, float* workspace, ae_stream stream);

/* Same, with the statistics taken from the fixed-point column sums that the GEMMs producing x1 / x2 accumulated in
 * their epilogues (ae_gemm_args.colstats, layout [B, C_i, 2] int64): one launch, one pass over the tensor. */
int ae_groupnorm_cs(const float* x1, int C1, const int64_t* colstats1, const float* x2, int C2, const int64_t* colstats2,
                    int B, int64_t HW, int groups, float eps, const float* gamma, const float* beta, int silu,
                    void* out_bf16, void* raw_out_bf16, float* cat_out_f32, float* workspace, ae_stream stream);

/* LayerNorm over the last dim of fp32 [rows, C] -> bf16 (attention.py:393-395, eps 1e-5) */
int ae_layernorm(const float* x, int64_t rows, int C, float eps, const float* gamma, const float* beta, void* out_bf16,
                 ae_stream stream);

/* GEGLU: out[m, j] = h[m, j] * gelu(h[m, inner + j])   (attention.py:37-44), bf16 in/out */
int ae_geglu(const void* h_bf16, int64_t rows, int inner, void* out_bf16, ae_stream stream);

/* Fused multi-head attention, online softmax (attention.py:285-323): O = softmax(Q K^T * scale + bias) V.
 * q: [B, Tq, heads*d] rows with stride ld_q (elements); k, v: [Bkv, Tk, heads*d] with strides ld_k, ld_v;
 * kv_batch_map[b] (optional, device) selects the K/V batch of query batch b (shared text K/V);
 * key_bias (optional): additive fp32 [Bkv, Tk] (models.py:204-210: 0 keep / -10000 discard).  bf16 in/out. */
int ae_attention(const void* q, int64_t ld_q, int64_t q_batch_stride, const void* k, int64_t ld_k,
                 int64_t k_batch_stride, const void* v, int64_t ld_v, int64_t v_batch_stride, const int32_t* kv_batch_map,
                 const float* key_bias, int64_t ld_bias, int B, int heads, int d, int Tq, int Tk, float scale, void* out,
                 int64_t ld_o, int64_t o_batch_stride, ae_stream stream);

/* ae_attention dispatch: sequences with Tq, Tk >= 128, d <= 128 (multiple of 8), no kv_batch_map and 16-byte aligned
 * strides run on the tcgen05 kernel (csrc/attn_tc.cu: S and the per-block P.V product in TMEM, Q / K / V by TMA, online
 * softmax from tcgen05.ld); the rest on the mma.sync kernel.  ae_set_attention_tc(0) forces the mma.sync kernel (A/B). */
void ae_set_attention_tc(int on);

/* Sinusoidal timestep embedding [cos | sin] (util.py:173-197), bf16 out [B, dim]; t: int64 [B] */
int ae_timestep_embedding(const int64_t* t, int B, int dim, void* out_bf16, ae_stream stream);

/* nearest-neighbour resize of channels-last fp32 [B,H,W,C] to [B,Ho,Wo,C] bf16 (openaimodel.py:113-121) */
int ae_upsample_nearest(const float* x, int B, int H, int W, int C, int Ho, int Wo, void* out_bf16, ae_stream stream);

/* layout / dtype movers */
int ae_nchw_to_nhwc(const float* x, int B, int C, int H, int W, float* out_f32, void* out_bf16, ae_stream stream);
int ae_nhwc_to_nchw(const float* x, int B, int C, int H, int W, float* out_f32, ae_stream stream);
int ae_cast_f32_bf16(const float* x, int64_t n, void* out_bf16, int silu, ae_stream stream);
int ae_add_f32(const float* a, const float* b, float scale_b, int64_t n, float* out, ae_stream stream);
/* row-wise softmax of fp32 [rows, n] (+ optional per-column bias) -> bf16, used by the unfused attention path (VAE) */
int ae_softmax_rows(const float* x, int64_t rows, int n, int64_t ld, void* out_bf16, int64_t ld_out, ae_stream stream);
/* bf16 [rows, cols] -> bf16 [cols, rows] per batch */
int ae_transpose_bf16(const void* x, int batch, int rows, int cols, void* out, ae_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Audio front end (code/audioldm/audio/stft.py:52-81,159-180; tools.py:18-31):
 * log-mel of a mono waveform: reflect pad n_fft/2, frames of n_fft with hop, window, |DFT|, mel matmul,
 * log(clamp(.,1e-5)).  wav [n_samples] f32; window [n_fft]; mel_basis [n_mels, n_fft/2+1]; out [n_frames, n_mels]
 * workspace: >= n_frames * (n_fft/2+1) floats (magnitudes). */
int ae_stft_mel(const float* wav, int n_samples, int n_fft, int hop, const float* window, const float* mel_basis,
                int n_mels, int n_frames, float* mag_workspace, float* out_logmel, ae_stream stream);

/* 1-D ops of the HiFi-GAN vocoder (code/audioldm/hifigan/models.py:20-165), channels-last [B,T,C] fp32 */
/* out = leaky_relu(scale * x, slope) as bf16 (scale carries the 1/num_kernels of the MRF average, models.py:160) */
int ae_leaky_relu_bf16(const float* x, int64_t n, float scale, float slope, void* out_bf16, ae_stream stream);
int ae_tanh_f32(const float* x, int64_t n, float* out, ae_stream stream);
/* waveform float [-1,1] -> 16-bit PCM, (x * 32768) truncated toward zero (hifigan/utilities.py:76-85 vocoder_infer) */
int ae_wave_to_int16(const float* x, int64_t n, int16_t* out, ae_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* AEDIT_H_ */
