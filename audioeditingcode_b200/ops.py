"""Device ops of the hot path: thin Python shims that hand raw device pointers of torch CUDA tensors to
libaedit.so (include/aedit.h) on torch's current stream.  torch is used only for memory and streams.
There is deliberately no torch/CPU implementation here — a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import AeGemmArgs, check

BF16 = torch.bfloat16
F32 = torch.float32


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.AeditError("libaedit ops need CUDA tensors (the hot path has no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class CudaOps:
    """The only ops backend of the product.  Tests inject a torch backend with the same surface to validate the
    executor's wiring on CPU; that backend lives under tests/ and is never importable from this package."""

    name = "cuda"

    SPLITK_WS_BYTES = 64 << 20

    def __init__(self):
        self.lib = _lib.load()
        self.act_dtype = _lib.operand_torch_dtype()      # fp16 (default build) or bf16 (AEDIT_OPERANDS=bf16)
        self._gn_ws = {}
        self._splitk_ws = {}
        import os
        # programmatic dependent launch: GEMM-only mode (weight prefetch ahead of the dependency wait) measured +2-4 %
        # end to end on B200, all-kernel mode measured slower (profiles/r01_bench_v7_*, r01_bench_v8_*)
        self.lib.ae_set_pdl(int(os.environ.get("AEDIT_PDL", "2")))   # 0 off, 1 all kernels, 2 GEMMs only
        if "AEDIT_PDL_EXTRA" in os.environ:
            self.lib.ae_set_pdl_extra(int(os.environ["AEDIT_PDL_EXTRA"]))
        if "AEDIT_GN_STREAM_MIN_BYTES" in os.environ:
            self.lib.ae_set_gn_stream_min_bytes(int(os.environ["AEDIT_GN_STREAM_MIN_BYTES"]))
        if "AEDIT_PERSIST_MIN_TILES" in os.environ:
            self.lib.ae_set_persistent_min_tiles(int(os.environ["AEDIT_PERSIST_MIN_TILES"]))
        if "AEDIT_ATTN_TC" in os.environ:
            self.lib.ae_set_attention_tc(int(os.environ["AEDIT_ATTN_TC"]))      # 0: mma.sync kernel everywhere (A/B)
        if "AEDIT_SHALLOW_KB" in os.environ:
            self.lib.ae_set_shallow_kblocks(int(os.environ["AEDIT_SHALLOW_KB"]))
        if "AEDIT_TILE_MODEL" in os.environ:
            self.lib.ae_set_tile_model(int(os.environ["AEDIT_TILE_MODEL"]))
        if "AEDIT_FAST_EPILOGUE" in os.environ:
            self.lib.ae_set_fast_epilogue(int(os.environ["AEDIT_FAST_EPILOGUE"]))
        if "AEDIT_SPLITK_CTAS" in os.environ:
            self.lib.ae_set_splitk_ctas(int(os.environ["AEDIT_SPLITK_CTAS"]))

    def _splitk_workspace(self, device):
        ws = self._splitk_ws.get(str(device))
        if ws is None:
            ws = torch.empty(self.SPLITK_WS_BYTES // 4, dtype=torch.float32, device=device)
            self._splitk_ws[str(device)] = ws
        return ws

    # ---------------------------------------------------------------- memory
    def empty(self, shape, dtype, device):
        return torch.empty(shape, dtype=dtype, device=device)

    def launch_count(self) -> int:
        return int(self.lib.ae_launch_count())

    # ---------------------------------------------------------------- GEMM / conv
    def gemm(self, A, W, *, out_f32=None, out_bf16=None, bias=None, rowbias=None, rows_per_group=1, residual=None,
             act=0, alpha=1.0, conv=None, M=None, K=None, force_bn=0, batch=1, strideA=0, strideW=0, stride_out=0,
             stride_res=0, lda=None, ldw=None, force_split=0, ld_out_f32=None, ld_out_bf16=None, force_stages=0,
             w_dynamic=False, colstats=None, cs_rows=0, force_persistent=0, softmax=None):
        """D = alpha*A@W^T (+bias)(+rowbias[row//rows_per_group])(+residual) -> act.  conv=(B,H,W,C,kh,kw,dh,dw)
        turns A (channels-last image) into an implicit-GEMM operand."""
        a = AeGemmArgs()
        a.A = A.data_ptr()
        a.W = W.data_ptr()
        N = W.shape[-2]
        if conv is not None:
            B_, H_, W_, C_, kh, kw, dh, dw = conv
            a.conv = 1
            a.B, a.H, a.W_, a.C, a.kh, a.kw, a.dil_h, a.dil_w = B_, H_, W_, C_, kh, kw, dh, dw
            a.M = B_ * H_ * W_
            a.K = kh * kw * C_
            a.lda = C_
        else:
            a.M = int(M if M is not None else A.shape[-2])
            a.K = int(K if K is not None else A.shape[-1])
            a.lda = int(lda if lda is not None else A.stride(-2))
        a.N = int(N)
        a.ldw = int(ldw if ldw is not None else W.stride(-2))
        a.batch = batch
        a.strideA, a.strideW, a.stride_out, a.stride_res = strideA, strideW, stride_out, stride_res
        if bias is not None:
            a.bias = bias.data_ptr()
        if rowbias is not None:
            a.rowbias = rowbias.data_ptr()
            a.ld_rowbias = rowbias.stride(0)
            a.rows_per_group = rows_per_group
        if residual is not None:
            a.residual = residual.data_ptr()
            a.ld_res = residual.stride(-2)
        if out_f32 is not None:
            a.out_f32 = out_f32.data_ptr()
            a.ld_out_f32 = int(ld_out_f32 if ld_out_f32 is not None else out_f32.stride(-2))
        if out_bf16 is not None:
            a.out_bf16 = out_bf16.data_ptr()
            a.ld_out_bf16 = int(ld_out_bf16 if ld_out_bf16 is not None else out_bf16.stride(-2))
        a.act = act
        a.alpha = alpha
        a.force_bn = force_bn
        a.force_split = force_split
        a.force_stages = force_stages
        a.w_dynamic = 1 if w_dynamic else 0
        a.force_persistent = force_persistent
        if softmax is not None:             # act 3: (keys per group, columns per text row, slot map, rows per sample, bias)
            L_, block, slot, rows, sbias = softmax
            a.act = 3
            a.sm_L, a.sm_block, a.sm_rows = int(L_), int(block), int(rows)
            a.sm_slot = slot.data_ptr()
            if sbias is not None:
                a.sm_bias = sbias.data_ptr()
        if colstats is not None:            # GroupNorm statistics of the output, accumulated by the epilogue
            a.colstats = colstats.data_ptr()
            a.cs_rows_per_sample = int(cs_rows)
        if batch == 1 and act != 2 and softmax is None:
            ws = self._splitk_workspace(A.device)
            a.splitk_ws = ws.data_ptr()
            a.splitk_ws_bytes = ws.numel() * 4
        check(self.lib.ae_gemm(C.byref(a), _stream()), "ae_gemm")

    def conv_supported(self, B, H, W, C_) -> bool:
        return bool(self.lib.ae_gemm_conv_supported(B, H, W, C_))

    def im2col(self, x, B, H, W, C_, kh, kw, stride, dil, pad_t, pad_l, Ho, Wo, out):
        check(self.lib.ae_im2col(_p(x), 0 if x.dtype == F32 else 1, B, H, W, C_, kh, kw, stride, dil, pad_t, pad_l,
                                 Ho, Wo, _p(out), out.stride(-2), _stream()), "ae_im2col")

    # ---------------------------------------------------------------- norms / activations
    def _gn_workspace(self, B, groups, device):
        key = (B, groups, str(device))
        ws = self._gn_ws.get(key)
        if ws is None:
            n = int(self.lib.ae_groupnorm_workspace_bytes(B, groups))
            ws = torch.zeros((n + 3) // 4, dtype=torch.float32, device=device)
            self._gn_ws[key] = ws
        return ws

    def groupnorm(self, x1, x2, gamma, beta, eps, groups, silu, out, raw_out=None, cat_out=None, cs1=None, cs2=None):
        """x1 [B,HW,C1] (+ x2 [B,HW,C2] virtually concatenated) fp32 -> out bf16 [B,HW,C1+C2].
        cs1 / cs2: int64 [B, C_i, 2] column statistics accumulated by the GEMMs that produced x1 / x2 (gemm(colstats=));
        with them the statistics pass over the tensor is skipped (one launch instead of two)."""
        B = x1.shape[0]
        C1 = x1.shape[-1]
        HW = x1.numel() // (B * C1)
        C2 = 0 if x2 is None else x2.shape[-1]
        ws = self._gn_workspace(B, groups, x1.device)
        if cs1 is not None and (x2 is None or cs2 is not None):
            check(self.lib.ae_groupnorm_cs(_p(x1), C1, _p(cs1), _p(x2), C2, _p(cs2), B, HW, groups, eps, _p(gamma),
                                           _p(beta), int(silu), _p(out), _p(raw_out), _p(cat_out), _p(ws), _stream()),
                  "ae_groupnorm_cs")
            return
        check(self.lib.ae_groupnorm(_p(x1), C1, _p(x2), C2, B, HW, groups, eps, _p(gamma), _p(beta), int(silu), _p(out),
                                    _p(raw_out), _p(cat_out), _p(ws), _stream()), "ae_groupnorm")

    def layernorm(self, x, gamma, beta, out, eps=1e-5):
        Cd = x.shape[-1]
        check(self.lib.ae_layernorm(_p(x), x.numel() // Cd, Cd, eps, _p(gamma), _p(beta), _p(out), _stream()),
              "ae_layernorm")

    def geglu(self, h, out):
        inner = out.shape[-1]
        check(self.lib.ae_geglu(_p(h), out.numel() // inner, inner, _p(out), _stream()), "ae_geglu")

    def attention(self, q, k, v, out, heads, d, scale, Tq, Tk, B, ld_q, bs_q, ld_k, bs_k, ld_v, bs_v, kv_map=None,
                  bias=None):
        check(self.lib.ae_attention(_p(q), ld_q, bs_q, _p(k), ld_k, bs_k, _p(v), ld_v, bs_v, _p(kv_map), _p(bias),
                                    0 if bias is None else bias.stride(0), B, heads, d, Tq, Tk, scale, _p(out),
                                    out.stride(-2), Tq * out.stride(-2), _stream()), "ae_attention")

    def timestep_embedding(self, t, dim, out):
        check(self.lib.ae_timestep_embedding(_p(t), t.shape[0], dim, _p(out), _stream()), "ae_timestep_embedding")

    def upsample_nearest(self, x, B, H, W, C_, Ho, Wo, out):
        check(self.lib.ae_upsample_nearest(_p(x), B, H, W, C_, Ho, Wo, _p(out), _stream()), "ae_upsample_nearest")

    def nchw_to_nhwc(self, x, out_f32=None, out_bf16=None):
        B, C_, H, W = x.shape
        check(self.lib.ae_nchw_to_nhwc(_p(x), B, C_, H, W, _p(out_f32), _p(out_bf16), _stream()), "ae_nchw_to_nhwc")

    def nhwc_to_nchw(self, x, B, C_, H, W, out):
        check(self.lib.ae_nhwc_to_nchw(_p(x), B, C_, H, W, _p(out), _stream()), "ae_nhwc_to_nchw")

    def cast_bf16(self, x, out, silu=False):
        check(self.lib.ae_cast_f32_bf16(_p(x), x.numel(), _p(out), int(silu), _stream()), "ae_cast_f32_bf16")

    def add(self, a, b, out, scale_b=1.0):
        check(self.lib.ae_add_f32(_p(a), _p(b), scale_b, a.numel(), _p(out), _stream()), "ae_add_f32")

    def softmax_rows(self, x, out):
        n = x.shape[-1]
        check(self.lib.ae_softmax_rows(_p(x), x.numel() // n, n, x.stride(-2), _p(out), out.stride(-2), _stream()),
              "ae_softmax_rows")

    def transpose_bf16(self, x, out):
        b = x.numel() // (x.shape[-1] * x.shape[-2])
        check(self.lib.ae_transpose_bf16(_p(x), b, x.shape[-2], x.shape[-1], _p(out), _stream()), "ae_transpose_bf16")

    def leaky_relu_bf16(self, x, slope, out, scale=1.0):
        check(self.lib.ae_leaky_relu_bf16(_p(x), x.numel(), scale, slope, _p(out), _stream()), "ae_leaky_relu_bf16")

    def tanh(self, x, out):
        check(self.lib.ae_tanh_f32(_p(x), x.numel(), _p(out), _stream()), "ae_tanh_f32")

    def wave_to_int16(self, x, out):
        check(self.lib.ae_wave_to_int16(_p(x), x.numel(), _p(out), _stream()), "ae_wave_to_int16")

    def stft_mel(self, wav, n_fft, hop, window, mel_basis, n_frames, out, mag_ws=None):
        check(self.lib.ae_stft_mel(_p(wav), wav.numel(), n_fft, hop, _p(window), _p(mel_basis), mel_basis.shape[0],
                                   n_frames, _p(mag_ws), _p(out), _stream()), "ae_stft_mel")
