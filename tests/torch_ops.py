"""TEST INFRASTRUCTURE — a torch (CPU, fp32) stand-in for audioeditingcode_b200.ops.CudaOps with the same
surface, used ONLY to validate the host-side wiring of UNetEngine (weight packing, layer order, skip/concat
bookkeeping, pointer-offset views) against the oracle on a machine without a GPU.  It is not importable from
the product package and is never used on a CUDA device."""
import math

import torch
import torch.nn.functional as F


class TorchOps:
    name = "torch-test"
    act_dtype = torch.float32

    def __init__(self):
        self.launches = 0

    def empty(self, shape, dtype, device):
        return torch.full(shape, float("nan"), dtype=dtype, device=device)

    def launch_count(self):
        return self.launches

    def conv_supported(self, B, H, W, C):
        if C % 64:
            return False
        if W >= 128:
            return W % 128 == 0
        if 128 % W:
            return False
        rows = 128 // W
        return (H % rows == 0) if H >= rows else (rows % H == 0)

    def gemm(self, A, W, *, out_f32=None, out_bf16=None, bias=None, rowbias=None, rows_per_group=1, residual=None,
             act=0, alpha=1.0, conv=None, M=None, K=None, force_bn=0, batch=1, strideA=0, strideW=0, stride_out=0,
             lda=None, ldw=None, ld_out_bf16=None, softmax=None, colstats=None, cs_rows=0, **kw):
        self.launches += 1
        if batch > 1:
            # independent problems z with element strides, exactly the pointer arithmetic of ae_gemm (include/aedit.h)
            N = W.shape[-2]
            Az = torch.as_strided(A, (batch, M, K), (strideA, lda, 1), A.storage_offset())
            Wz = torch.as_strided(W, (batch, N, K), (strideW, ldw, 1), W.storage_offset())
            Oz = torch.as_strided(out_bf16, (batch, M, N), (stride_out, ld_out_bf16, 1), out_bf16.storage_offset())
            Oz.copy_(alpha * torch.einsum("zmk,znk->zmn", Az, Wz))
            return
        if softmax is not None:
            # act 3 (include/aedit.h, ae_gemm_args.sm_*): per (text row, head) softmax over L columns for the sample's
            # own text row, zeros for the other text rows
            L, block, slot, rows, sbias = softmax
            acc = A.reshape(-1, A.shape[-1]) @ W.t()
            Mr, N = acc.shape
            r_of_row = slot.long()[torch.arange(Mr) // rows]                        # [M]
            g = acc.reshape(Mr, N // L, L)
            r_of_group = (torch.arange(N // L) * L) // block                        # [N/L]
            if sbias is not None:
                g = g + sbias[r_of_group][None]
            p = torch.softmax(g, dim=-1)
            p = p * (r_of_group[None, :] == r_of_row[:, None]).to(p.dtype)[..., None]
            out_bf16.copy_(p.reshape(Mr, N))
            return
        if conv is not None:
            B, H, W_, C, kh, kw_, dh, dw = conv
            x = A.reshape(B, H, W_, C).permute(0, 3, 1, 2)
            wt = W.reshape(W.shape[0], kh, kw_, C).permute(0, 3, 1, 2)
            y = F.conv2d(x, wt, padding=(dh * (kh - 1) // 2, dw * (kw_ - 1) // 2), dilation=(dh, dw))
            acc = y.permute(0, 2, 3, 1).reshape(B * H * W_, -1)
        else:
            K = K if K is not None else A.shape[-1]
            A2 = A.reshape(-1, A.shape[-1])[:, :K]
            acc = A2 @ W[:, :K].t()
        acc = acc * alpha
        if bias is not None:
            acc = acc + bias
        if rowbias is not None:
            idx = torch.arange(acc.shape[0]) // rows_per_group
            acc = acc + rowbias[idx]
        if residual is not None:
            acc = acc + residual.reshape(acc.shape)
        if act == 1:
            acc = F.silu(acc)
        elif act == 2:
            a3 = acc.reshape(acc.shape[0], -1, 2, 16)
            acc = (a3[:, :, 0] * F.gelu(a3[:, :, 1])).reshape(acc.shape[0], -1)
        if out_f32 is not None:
            out_f32.reshape(acc.shape).copy_(acc) if out_f32.is_contiguous() else out_f32.copy_(acc)
        if out_bf16 is not None:
            out_bf16.copy_(acc)
        if colstats is not None:
            # ae_gemm_args.colstats: fixed-point per-(sample, column) sum / sum of squares, ACCUMULATED into the caller's
            # zeroed int64 buffer [sample][N][2]
            xs = acc.double().reshape(-1, cs_rows, acc.shape[-1])
            add = torch.stack([(xs.sum(1) * 2.0 ** 28).round(), ((xs * xs).sum(1) * 2.0 ** 24).round()], -1).to(torch.int64)
            colstats.add_(add.reshape(-1))

    def im2col(self, x, B, H, W, C, kh, kw, stride, dil, pad_t, pad_l, Ho, Wo, out):
        self.launches += 1
        xi = x.reshape(B, H, W, C).permute(0, 3, 1, 2).float()
        pad_b = max(0, (Ho - 1) * stride + (kh - 1) * dil + 1 - H - pad_t)
        pad_r = max(0, (Wo - 1) * stride + (kw - 1) * dil + 1 - W - pad_l)
        xp = F.pad(xi, (pad_l, pad_r, pad_t, pad_b))
        cols = F.unfold(xp, (kh, kw), dilation=dil, stride=stride)          # [B, C*kh*kw, L]
        cols = cols.reshape(B, C, kh * kw, -1)[:, :, :, : Ho * Wo]
        cols = cols.permute(0, 3, 2, 1).reshape(B * Ho * Wo, kh * kw * C)    # (tap, c) ordering
        out.zero_()
        out[:, : kh * kw * C] = cols

    def groupnorm(self, x1, x2, gamma, beta, eps, groups, silu, out, raw_out=None, cat_out=None, cs1=None, cs2=None):
        self.launches += 1
        x = x1 if x2 is None else torch.cat([x1, x2], dim=-1)
        B, C = x.shape[0], x.shape[-1]
        xc = x.reshape(B, -1, C).permute(0, 2, 1)
        if cs1 is not None:
            # ae_groupnorm_cs: statistics from the producers' fixed-point column sums (concatenated inputs: two buffers)
            cs = cs1.reshape(B, -1, 2) if cs2 is None else torch.cat([cs1.reshape(B, -1, 2), cs2.reshape(B, -1, 2)], 1)
            n = xc.shape[-1] * (C // groups)
            su = cs[..., 0].reshape(B, groups, -1).sum(-1).double() / 2.0 ** 28
            sq = cs[..., 1].reshape(B, groups, -1).sum(-1).double() / 2.0 ** 24
            mean = su / n
            rstd = 1.0 / torch.sqrt((sq / n - mean * mean).clamp_min(0).float() + eps)
            mean_c = mean.float().repeat_interleave(C // groups, 1)[:, :, None]
            rstd_c = rstd.repeat_interleave(C // groups, 1)[:, :, None]
            y = ((xc - mean_c) * rstd_c * gamma[None, :, None] + beta[None, :, None]).permute(0, 2, 1)
        else:
            y = F.group_norm(xc, groups, gamma, beta, eps).permute(0, 2, 1)
        if silu:
            y = F.silu(y)
        out.reshape(B, -1, C).copy_(y)
        if raw_out is not None:
            raw_out.reshape(B, -1, C).copy_(x.reshape(B, -1, C))
        if cat_out is not None:
            cat_out.reshape(B, -1, C).copy_(x.reshape(B, -1, C))

    def layernorm(self, x, gamma, beta, out, eps=1e-5):
        self.launches += 1
        out.copy_(F.layer_norm(x, x.shape[-1:], gamma, beta, eps))

    def geglu(self, h, out):
        self.launches += 1
        a, g = h.chunk(2, dim=-1)
        out.copy_(a * F.gelu(g))

    def attention(self, q, k, v, out, heads, d, scale, Tq, Tk, B, ld_q, bs_q, ld_k, bs_k, ld_v, bs_v, kv_map=None,
                  bias=None):
        self.launches += 1
        C = heads * d
        q2 = q.reshape(-1, q.shape[-1])[:, :C].reshape(B, Tq, heads, d).transpose(1, 2)
        k2 = k.reshape(-1, k.shape[-1])[:, :C].reshape(-1, Tk, heads, d).transpose(1, 2)
        v2 = v.reshape(-1, v.shape[-1])[:, :C].reshape(-1, Tk, heads, d).transpose(1, 2)
        if kv_map is not None:
            k2, v2 = k2[kv_map.long()], v2[kv_map.long()]
        s = q2 @ k2.transpose(-1, -2) * scale
        if bias is not None:
            bb = bias if kv_map is None else bias[kv_map.long()]
            s = s + bb[:, None, None, :]
        o = s.softmax(-1) @ v2
        out.copy_(o.transpose(1, 2).reshape(B * Tq, C))

    def timestep_embedding(self, t, dim, out):
        self.launches += 1
        half = dim // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
        args = t[:, None].float() * freqs[None]
        out.copy_(torch.cat([torch.cos(args), torch.sin(args)], -1))

    def upsample_nearest(self, x, B, H, W, C, Ho, Wo, out):
        self.launches += 1
        xi = x.reshape(B, H, W, C).permute(0, 3, 1, 2)
        out.copy_(F.interpolate(xi, size=(Ho, Wo), mode="nearest").permute(0, 2, 3, 1))

    def nchw_to_nhwc(self, x, out_f32=None, out_bf16=None):
        self.launches += 1
        y = x.permute(0, 2, 3, 1)
        if out_f32 is not None:
            out_f32.copy_(y)
        if out_bf16 is not None:
            out_bf16.copy_(y)

    def nhwc_to_nchw(self, x, B, C, H, W, out):
        self.launches += 1
        out.copy_(x.reshape(B, H, W, C).permute(0, 3, 1, 2))

    def cast_bf16(self, x, out, silu=False):
        self.launches += 1
        out.copy_(F.silu(x) if silu else x)

    def add(self, a, b, out, scale_b=1.0):
        self.launches += 1
        out.copy_(a + scale_b * b)
