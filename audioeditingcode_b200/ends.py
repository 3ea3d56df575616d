"""The audio "ends" of the editing path on the same device kernels as the U-Net:

  VAEEngine      models.py:495-503  `vae.encode(x).latent_dist.mode() * scaling_factor` / `vae.decode(z / scaling_factor)`
                 ([UPSTREAM] diffusers AutoencoderKL; in-tree statement code/audioldm/variational_autoencoder/modules.py:419-683)
  HiFiGANEngine  models.py:505-509  `mel_spectrogram_to_waveform`
                 ([UPSTREAM] transformers SpeechT5HifiGan; in-tree statement code/audioldm/hifigan/models.py:20-165)

Convolutions are tcgen05 implicit GEMMs (2-D for the VAE, 1-D dilated for the vocoder: the same 4-D TMA box with
H = 1), transposed convolutions are decomposed into `stride` phase GEMMs over a shared patch matrix, GroupNorm /
activations / the single-head mid-block attention use the library kernels.  Weights use diffusers / transformers
state-dict names; without a checkpoint directory seeded synthetic weights of the same architecture are used
(no network in the build / bench environment).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Tuple

import torch

from .ops import CudaOps

F32 = torch.float32

VAE_CH, VAE_MULT, VAE_RES, VAE_Z = 128, (1, 2, 4), 2, 8
VAE_SCALING = 0.9227914214134216
HIFI_RATES, HIFI_KERNELS, HIFI_INIT = (5, 4, 2, 2, 2), (16, 16, 8, 4, 4), 1024
HIFI_RES_K, HIFI_RES_D, HIFI_MELS = (3, 7, 11), ((1, 3, 5), (1, 3, 5), (1, 3, 5)), 64


# ------------------------------------------------------------------------------------------------- weights
def vae_weight_shapes() -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    def res(p, ci, co):
        norm(p + ".norm1", ci)
        conv(p + ".conv1", co, ci, 3)
        norm(p + ".norm2", co)
        conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1)

    def mid(p, c):
        res(p + ".resnets.0", c, c)
        a = p + ".attentions.0"
        norm(a + ".group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{a}.{n}", c, c)
        res(p + ".resnets.1", c, c)

    chs = [VAE_CH * m for m in VAE_MULT]
    conv("encoder.conv_in", chs[0], 1, 3)
    c = chs[0]
    for i, co in enumerate(chs):
        for j in range(VAE_RES):
            res(f"encoder.down_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i != len(chs) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
    mid("encoder.mid_block", c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", 2 * VAE_Z, c, 3)
    conv("quant_conv", 2 * VAE_Z, 2 * VAE_Z, 1)
    conv("post_quant_conv", VAE_Z, VAE_Z, 1)
    conv("decoder.conv_in", c, VAE_Z, 3)
    mid("decoder.mid_block", c)
    for i, co in enumerate(reversed(chs)):
        for j in range(VAE_RES + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i != len(chs) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("decoder.conv_norm_out", c)
    conv("decoder.conv_out", 1, c, 3)
    return s


def hifigan_weight_shapes() -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {"conv_pre.weight": (HIFI_INIT, HIFI_MELS, 7), "conv_pre.bias": (HIFI_INIT,)}
    ch = HIFI_INIT
    for i, (u, k) in enumerate(zip(HIFI_RATES, HIFI_KERNELS)):
        s[f"upsampler.{i}.weight"] = (ch, ch // 2, k)
        s[f"upsampler.{i}.bias"] = (ch // 2,)
        ch //= 2
        for j, (rk, rd) in enumerate(zip(HIFI_RES_K, HIFI_RES_D)):
            for d in range(len(rd)):
                for c in ("convs1", "convs2"):
                    s[f"resblocks.{i * 3 + j}.{c}.{d}.weight"] = (ch, ch, rk)
                    s[f"resblocks.{i * 3 + j}.{c}.{d}.bias"] = (ch,)
    s["conv_post.weight"] = (1, ch, 7)
    s["conv_post.bias"] = (1,)
    return s


def synthetic(shapes: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith(".weight") and len(shp) == 1:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias") or name in ("mean",):
            t = 0.02 * torch.randn(shp, generator=g)
        elif name == "scale":
            t = 1.0 + 0.1 * torch.rand(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) * (0.7 / math.sqrt(fan_in))
        out[name] = t
    return out


def _load_state(path: str) -> Dict[str, torch.Tensor]:
    for f in ("diffusion_pytorch_model.safetensors", "model.safetensors"):
        fp = os.path.join(path, f)
        if os.path.exists(fp):
            from safetensors.torch import load_file
            return {k: v.float() for k, v in load_file(fp).items()}
    for f in ("diffusion_pytorch_model.bin", "pytorch_model.bin"):
        fp = os.path.join(path, f)
        if os.path.exists(fp):
            return {k: v.float() for k, v in torch.load(fp, map_location="cpu", weights_only=True).items()}
    raise FileNotFoundError(path)


def ldm_vae_to_canonical(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """State dict of the ORIGINAL AudioLDM `AutoencoderKL` (what a TANGO snapshot's pytorch_model_vae.bin holds,
    models.py:410-421; in-tree module: code/audioldm/variational_autoencoder/modules.py:419-683, autoencoder.py:9-103)
    -> the diffusers names this package uses.  LDM `down.{i}.block.{j}` / `up.{level}.block.{j}` (level counted from the
    full resolution, the decoder walks it downwards) / `mid.block_1|attn_1|block_2` / `nin_shortcut` / 1x1-conv attention
    projections `q k v proj_out` -> `down_blocks` / `up_blocks.{n-1-level}` / `mid_block.resnets|attentions` /
    `conv_shortcut` / Linear `to_q to_k to_v to_out.0`."""
    n = len(VAE_MULT)
    out: Dict[str, torch.Tensor] = {}

    def res_names(dst, src):
        return [(f"{src}.{b}", f"{dst}.{a}") for a, b in (("norm1", "norm1"), ("conv1", "conv1"), ("norm2", "norm2"),
                                                          ("conv2", "conv2"), ("conv_shortcut", "nin_shortcut"))]
    pairs = []
    for side in ("encoder", "decoder"):
        pairs += [(f"{side}.conv_in", f"{side}.conv_in"), (f"{side}.norm_out", f"{side}.conv_norm_out"),
                  (f"{side}.conv_out", f"{side}.conv_out")]
        pairs += res_names(f"{side}.mid_block.resnets.0", f"{side}.mid.block_1")
        pairs += res_names(f"{side}.mid_block.resnets.1", f"{side}.mid.block_2")
        pairs += [(f"{side}.mid.attn_1.norm", f"{side}.mid_block.attentions.0.group_norm")]
        pairs += [(f"{side}.mid.attn_1.{b}", f"{side}.mid_block.attentions.0.{a}")
                  for a, b in (("to_q", "q"), ("to_k", "k"), ("to_v", "v"), ("to_out.0", "proj_out"))]
    for i in range(n):
        for j in range(VAE_RES):
            pairs += res_names(f"encoder.down_blocks.{i}.resnets.{j}", f"encoder.down.{i}.block.{j}")
        pairs.append((f"encoder.down.{i}.downsample.conv", f"encoder.down_blocks.{i}.downsamplers.0.conv"))
        for j in range(VAE_RES + 1):
            pairs += res_names(f"decoder.up_blocks.{i}.resnets.{j}", f"decoder.up.{n - 1 - i}.block.{j}")
        pairs.append((f"decoder.up.{n - 1 - i}.upsample.conv", f"decoder.up_blocks.{i}.upsamplers.0.conv"))
    pairs += [("quant_conv", "quant_conv"), ("post_quant_conv", "post_quant_conv")]
    for src, dst in pairs:
        for suf in (".weight", ".bias"):
            t = sd.get(src + suf)
            if t is None:
                continue
            if ".attentions.0.to_" in dst and suf == ".weight" and t.dim() == 4:
                t = t[:, :, 0, 0]                                   # 1x1 conv -> Linear
            out[dst + suf] = t.float()
    return out


def ldm_hifigan_to_canonical(sd: Dict[str, torch.Tensor], prefix: str = "vocoder.") -> Dict[str, torch.Tensor]:
    """`vocoder.*` entries of the same checkpoint (code/audioldm/hifigan/models.py:112-165, weight norm already removed by
    get_vocoder, utilities.py:67-72) -> transformers SpeechT5HifiGan names (`ups.` -> `upsampler.`)."""
    return {k[len(prefix):].replace("ups.", "upsampler."): v.float() for k, v in sd.items() if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------------- shared helpers
class _Net:
    def __init__(self, device, ops: Optional[CudaOps] = None):
        self.device = torch.device(device)
        self.ops = ops or CudaOps()
        self.adt = self.ops.act_dtype
        self.w: Dict[str, torch.Tensor] = {}

    def _to(self, t, dtype):
        return t.detach().to(device=self.device, dtype=dtype).contiguous()

    def _conv2d(self, a_bf16, B, H, W, Cin, name, out, residual=None, k=3):
        """stride-1 'same' conv on a channels-last bf16 operand; implicit TMA gather when the geometry allows."""
        ops = self.ops
        Wt, bias = self.w[name + ".weight"], self.w[name + ".bias"]
        if k == 1:
            ops.gemm(a_bf16.reshape(B * H * W, Cin), Wt, out_f32=out, bias=bias, residual=residual)
        elif W >= 4 and ops.conv_supported(B, H, W, Cin):
            ops.gemm(a_bf16, Wt, out_f32=out, bias=bias, residual=residual, conv=(B, H, W, Cin, k, k, 1, 1))
        else:
            K = k * k * Cin
            col = ops.empty((B * H * W, (K + 7) // 8 * 8), self.adt, self.device)
            ops.im2col(a_bf16, B, H, W, Cin, k, k, 1, 1, k // 2, k // 2, H, W, col)
            ops.gemm(col, Wt, out_f32=out, bias=bias, residual=residual, K=K)


# ------------------------------------------------------------------------------------------------- VAE
class VAEEngine(_Net):
    def __init__(self, device, weights: Optional[Dict[str, torch.Tensor]] = None, scaling_factor: float = VAE_SCALING,
                 ops=None, seed: int = 0):
        super().__init__(device, ops)
        self.scaling = scaling_factor
        w = weights if weights is not None else synthetic(vae_weight_shapes(), seed)
        for name, t in w.items():
            if name.endswith(".weight") and t.dim() == 4:
                w2 = t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)
                if w2.shape[1] % 8:                                # TMA row pitch must be a multiple of 16 bytes
                    w2 = torch.nn.functional.pad(w2, (0, 8 - w2.shape[1] % 8))
                self.w[name] = self._to(w2, self.adt)
            elif name.endswith(".weight") and t.dim() == 2:
                self.w[name] = self._to(t, self.adt)
            else:
                self.w[name] = self._to(t, F32)
        for p in ("encoder.mid_block.attentions.0", "decoder.mid_block.attentions.0"):
            self.w[p + ".qk.weight"] = self._to(torch.cat([w[p + ".to_q.weight"], w[p + ".to_k.weight"]], 0), self.adt)
            self.w[p + ".qk.bias"] = self._to(torch.cat([w[p + ".to_q.bias"], w[p + ".to_k.bias"]], 0), F32)
        # latent_dist.mode() * scaling_factor: keep the mean rows of quant_conv, fold the scale (models.py:499)
        qw, qb = w["quant_conv.weight"].reshape(2 * VAE_Z, -1), w["quant_conv.bias"]
        self.w["quant_mean.weight"] = self._to(qw[:VAE_Z] * scaling_factor, self.adt)
        self.w["quant_mean.bias"] = self._to(qb[:VAE_Z] * scaling_factor, F32)
        # decode(z / scaling_factor): fold 1/scale into post_quant_conv (models.py:503)
        self.w["post_quant_scaled.weight"] = self._to(w["post_quant_conv.weight"].reshape(VAE_Z, -1) / scaling_factor,
                                                      self.adt)

    def _resnet(self, x, B, H, W, p):
        ops = self.ops
        Cin = x.shape[-1]
        Cout = self.w[p + ".conv1.bias"].shape[0]
        M = B * H * W
        has_sc = (p + ".conv_shortcut.weight") in self.w
        a1 = ops.empty((B, H, W, Cin), self.adt, self.device)
        raw = ops.empty((M, Cin), self.adt, self.device) if has_sc else None
        ops.groupnorm(x, None, self.w[p + ".norm1.weight"], self.w[p + ".norm1.bias"], 1e-6, 32, True, a1, raw_out=raw)
        h = ops.empty((M, Cout), F32, self.device)
        self._conv2d(a1, B, H, W, Cin, p + ".conv1", h)
        a2 = ops.empty((B, H, W, Cout), self.adt, self.device)
        ops.groupnorm(h.view(B, H * W, Cout), None, self.w[p + ".norm2.weight"], self.w[p + ".norm2.bias"], 1e-6, 32,
                      True, a2)
        if has_sc:
            res = ops.empty((M, Cout), F32, self.device)
            ops.gemm(raw, self.w[p + ".conv_shortcut.weight"], out_f32=res, bias=self.w[p + ".conv_shortcut.bias"])
        else:
            res = x.reshape(M, Cout)
        out = ops.empty((M, Cout), F32, self.device)
        self._conv2d(a2, B, H, W, Cout, p + ".conv2", out, residual=res)
        return out.view(B, H * W, Cout)

    def _attn(self, x, B, T, p):
        """single-head attention over T tokens, head dim = C = 512 (modules.py:203-230): unfused GEMM path."""
        ops = self.ops
        C = x.shape[-1]
        M = B * T
        g = ops.empty((M, C), self.adt, self.device)
        ops.groupnorm(x, None, self.w[p + ".group_norm.weight"], self.w[p + ".group_norm.bias"], 1e-6, 32, False, g)
        qk = ops.empty((M, 2 * C), self.adt, self.device)
        ops.gemm(g, self.w[p + ".qk.weight"], out_bf16=qk, bias=self.w[p + ".qk.bias"])
        v = ops.empty((M, C), self.adt, self.device)
        ops.gemm(g, self.w[p + ".to_v.weight"], out_bf16=v, bias=self.w[p + ".to_v.bias"])
        vT = ops.empty((B, C, T), self.adt, self.device)
        ops.transpose_bf16(v.view(B, T, C), vT)
        a = ops.empty((M, C), self.adt, self.device)
        for b in range(B):
            s = ops.empty((T, T), F32, self.device)
            qb = qk[b * T:(b + 1) * T]
            ops.gemm(qb, qb[:, C:], out_f32=s, alpha=float(C) ** -0.5, K=C, lda=2 * C, ldw=2 * C, force_split=1,
                     w_dynamic=True)
            pm = ops.empty((T, T), self.adt, self.device)
            ops.softmax_rows(s, pm)
            ops.gemm(pm, vT[b], out_bf16=a[b * T:(b + 1) * T], force_split=1, w_dynamic=True)
        out = ops.empty((M, C), F32, self.device)
        ops.gemm(a, self.w[p + ".to_out.0.weight"], out_f32=out, bias=self.w[p + ".to_out.0.bias"],
                 residual=x.reshape(M, C))
        return out.view(B, T, C)

    def _mid(self, x, B, H, W, p):
        x = self._resnet(x, B, H, W, p + ".resnets.0")
        x = self._attn(x, B, H * W, p + ".attentions.0")
        return self._resnet(x, B, H, W, p + ".resnets.1")

    def _encode_trunk(self, x):
        ops = self.ops
        x = x.to(self.device, F32).contiguous()
        B, _, H, W = x.shape
        xin = x.reshape(B, H, W, 1)                               # C = 1: NCHW == NHWC
        col = ops.empty((B * H * W, 16), self.adt, self.device)
        ops.im2col(xin, B, H, W, 1, 3, 3, 1, 1, 1, 1, H, W, col)
        h = ops.empty((B * H * W, VAE_CH), F32, self.device)
        ops.gemm(col, self.w["encoder.conv_in.weight"], out_f32=h, bias=self.w["encoder.conv_in.bias"])  # K padded 9->16
        h = h.view(B, H * W, VAE_CH)
        n = len(VAE_MULT)
        for i in range(n):
            for j in range(VAE_RES):
                h = self._resnet(h, B, H, W, f"encoder.down_blocks.{i}.resnets.{j}")
            if i != n - 1:
                C = h.shape[-1]
                Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1        # pad (0,1,0,1) then k3 s2 p0 (modules.py:87-89)
                col = ops.empty((B * Ho * Wo, 9 * C), self.adt, self.device)
                ops.im2col(h, B, H, W, C, 3, 3, 2, 1, 0, 0, Ho, Wo, col)
                d = ops.empty((B * Ho * Wo, C), F32, self.device)
                p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                ops.gemm(col, self.w[p + ".weight"], out_f32=d, bias=self.w[p + ".bias"])
                H, W = Ho, Wo
                h = d.view(B, H * W, C)
        h = self._mid(h, B, H, W, "encoder.mid_block")
        C = h.shape[-1]
        a = ops.empty((B, H, W, C), self.adt, self.device)
        ops.groupnorm(h, None, self.w["encoder.conv_norm_out.weight"], self.w["encoder.conv_norm_out.bias"], 1e-6, 32,
                      True, a)
        mom_in = ops.empty((B * H * W, 2 * VAE_Z), F32, self.device)
        self._conv2d(a, B, H, W, C, "encoder.conv_out", mom_in)
        mb = ops.empty((B * H * W, 2 * VAE_Z), self.adt, self.device)
        ops.cast_bf16(mom_in, mb)
        return mb, B, H, W

    def encode_mode(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B,1,T,64] log-mel -> latent mean * scaling_factor, NCHW [B,8,T/4,16] (models.py:495-499)."""
        ops = self.ops
        mb, B, H, W = self._encode_trunk(x)
        z = ops.empty((B * H * W, VAE_Z), F32, self.device)
        ops.gemm(mb, self.w["quant_mean.weight"], out_f32=z, bias=self.w["quant_mean.bias"])
        out = ops.empty((B, VAE_Z, H, W), F32, self.device)
        ops.nhwc_to_nchw(z, B, VAE_Z, H, W, out)
        return out

    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """Full moments [B,16,T/4,16] (mean | logvar), unscaled — autoencoder.py:49-55."""
        ops = self.ops
        mb, B, H, W = self._encode_trunk(x)
        m = ops.empty((B * H * W, 2 * VAE_Z), F32, self.device)
        ops.gemm(mb, self.w["quant_conv.weight"], out_f32=m, bias=self.w["quant_conv.bias"])
        out = ops.empty((B, 2 * VAE_Z, H, W), F32, self.device)
        ops.nhwc_to_nchw(m, B, 2 * VAE_Z, H, W, out)
        return out

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z: NCHW [B,8,h,w] (scaled latent) -> log-mel [B,1,4h,4w] (models.py:502-503)."""
        ops = self.ops
        z = z.to(self.device, F32).contiguous()
        B, Cz, H, W = z.shape
        zb = ops.empty((B, H, W, Cz), self.adt, self.device)
        ops.nchw_to_nhwc(z, out_bf16=zb)
        pq = ops.empty((B * H * W, Cz), F32, self.device)
        ops.gemm(zb.reshape(-1, Cz), self.w["post_quant_scaled.weight"], out_f32=pq, bias=self.w["post_quant_conv.bias"])
        C = VAE_CH * VAE_MULT[-1]
        col = ops.empty((B * H * W, 9 * Cz), self.adt, self.device)
        ops.im2col(pq.view(B, H, W, Cz), B, H, W, Cz, 3, 3, 1, 1, 1, 1, H, W, col)
        h = ops.empty((B * H * W, C), F32, self.device)
        ops.gemm(col, self.w["decoder.conv_in.weight"], out_f32=h, bias=self.w["decoder.conv_in.bias"])
        h = self._mid(h.view(B, H * W, C), B, H, W, "decoder.mid_block")
        n = len(VAE_MULT)
        for i in range(n):
            for j in range(VAE_RES + 1):
                h = self._resnet(h, B, H, W, f"decoder.up_blocks.{i}.resnets.{j}")
            if i != n - 1:
                C = h.shape[-1]
                up = ops.empty((B, 2 * H, 2 * W, C), self.adt, self.device)
                ops.upsample_nearest(h, B, H, W, C, 2 * H, 2 * W, up)
                H, W = 2 * H, 2 * W
                u = ops.empty((B * H * W, C), F32, self.device)
                self._conv2d(up, B, H, W, C, f"decoder.up_blocks.{i}.upsamplers.0.conv", u)
                h = u.view(B, H * W, C)
        C = h.shape[-1]
        a = ops.empty((B, H, W, C), self.adt, self.device)
        ops.groupnorm(h, None, self.w["decoder.conv_norm_out.weight"], self.w["decoder.conv_norm_out.bias"], 1e-6, 32,
                      True, a)
        o = ops.empty((B * H * W, 1), F32, self.device)
        self._conv2d(a, B, H, W, C, "decoder.conv_out", o)
        return o.view(B, 1, H, W)                                   # single channel: NHWC == NCHW


# ------------------------------------------------------------------------------------------------- HiFi-GAN
class HiFiGANEngine(_Net):
    def __init__(self, device, weights: Optional[Dict[str, torch.Tensor]] = None, ops=None, seed: int = 0,
                 normalize_before: bool = False):
        super().__init__(device, ops)
        w = weights if weights is not None else synthetic(hifigan_weight_shapes(), seed)
        self.normalize_before = normalize_before and "mean" in w
        if self.normalize_before:
            self.mean, self.scale = self._to(w["mean"], F32), self._to(w["scale"], F32)
        for name, t in w.items():
            if name.startswith("upsampler.") or name in ("mean", "scale"):
                continue
            if name.endswith(".weight"):                          # Conv1d [O, I, k] -> [O, k*I]  (tap, channel) order
                self.w[name] = self._to(t.permute(0, 2, 1).reshape(t.shape[0], -1), self.adt)
            else:
                self.w[name] = self._to(t, F32)
        # ConvTranspose1d [Cin, Cout, k], stride u, padding (k-u)//2 -> u phase matrices over ceil(k/u) taps
        self.up = []
        for i, (u, k) in enumerate(zip(HIFI_RATES, HIFI_KERNELS)):
            wt = w[f"upsampler.{i}.weight"]
            cin, cout = wt.shape[0], wt.shape[1]
            nt = (k + u - 1) // u
            phases = []
            for r in range(u):
                Wr = torch.zeros(cout, nt, cin)
                for jj in range(nt):
                    j = r + (nt - 1 - jj) * u
                    if j < k:
                        Wr[:, jj, :] = wt[:, :, j].t()
                phases.append(self._to(Wr.reshape(cout, nt * cin), self.adt))
            self.up.append(dict(u=u, k=k, pad=(k - u) // 2, nt=nt, cin=cin, cout=cout, phases=phases,
                                bias=self._to(w[f"upsampler.{i}.bias"], F32)))

    def _conv1d(self, a_bf16, B, T, C, name, k, dil, out, residual=None):
        """'same' dilated conv over time on a channels-last operand [B, T, C] (zero padding per clip: the implicit-conv
        TMA box never crosses the batch axis)."""
        ops = self.ops
        Wt, bias = self.w[name + ".weight"], self.w[name + ".bias"]
        if ops.conv_supported(B, 1, T, C):
            ops.gemm(a_bf16, Wt, out_f32=out, bias=bias, residual=residual, conv=(B, 1, T, C, 1, k, 1, dil))
        else:
            K = k * C
            col = ops.empty((B * T, (K + 7) // 8 * 8), self.adt, self.device)
            ops.im2col(a_bf16, B, 1, T, C, 1, k, 1, dil, 0, dil * (k - 1) // 2, 1, T, col)
            ops.gemm(col, Wt, out_f32=out, bias=bias, residual=residual, K=K)

    def _many(self, mel: torch.Tensor) -> torch.Tensor:
        """mel: [B, T, 64] log-mel -> waveforms [B, T_out] (hifigan/models.py:147-165), the B clips in ONE launch sequence
        (main_run.py:184-185 vocodes the edited and the original spectrogram: decode_to_mel on a 2-row batch halves the
        launches and doubles every GEMM's M)."""
        ops = self.ops
        B, T = mel.shape[0], mel.shape[1]
        x0 = mel.to(self.device, F32).contiguous()
        if self.normalize_before:
            x0 = ((x0 - self.mean) / self.scale).contiguous()
        xb = ops.empty((B, T, HIFI_MELS), self.adt, self.device)
        ops.cast_bf16(x0, xb)
        x = ops.empty((B * T, HIFI_INIT), F32, self.device)
        self._conv1d(xb, B, T, HIFI_MELS, "conv_pre", 7, 1, x)
        scale = 1.0
        for i, st in enumerate(self.up):
            u, pad, nt, cin, cout = st["u"], st["pad"], st["nt"], st["cin"], st["cout"]
            a = ops.empty((B, T, cin), self.adt, self.device)
            ops.leaky_relu_bf16(x, 0.1, a, scale=scale)            # scale: the /3 of the previous MRF average
            T_out = (T - 1) * u - 2 * pad + st["k"]
            Q = (T_out - 1 + pad) // u + 1
            col2 = ops.empty((B * Q, nt * cin), self.adt, self.device)     # 2-D: im2col takes the ROW pitch from stride(0)
            ops.im2col(a, B, 1, T, cin, 1, nt, 1, 1, 0, nt - 1, 1, Q, col2)
            col = col2.view(B, Q, nt * cin)
            y = ops.empty((B, T_out, cout), F32, self.device)
            for r in range(u):
                t0 = (r - pad) % u                                  # first output index of this phase
                q0 = (t0 + pad) // u
                cnt = (T_out - t0 + u - 1) // u
                if cnt <= 0:
                    continue
                for b in range(B):                                  # phase GEMMs interleave rows in time: one per clip
                    ops.gemm(col[b, q0:q0 + cnt], st["phases"][r], out_f32=y[b, t0:], bias=st["bias"], M=cnt,
                             ld_out_f32=u * cout)
            T = T_out
            xs = None
            for j, (rk, rd) in enumerate(zip(HIFI_RES_K, HIFI_RES_D)):
                p = f"resblocks.{i * 3 + j}"
                cur = y.view(B * T, cout)
                for d, dil in enumerate(rd):                        # hifigan/models.py:96-103
                    a1 = ops.empty((B, T, cout), self.adt, self.device)
                    ops.leaky_relu_bf16(cur, 0.1, a1)
                    h = ops.empty((B * T, cout), F32, self.device)
                    self._conv1d(a1, B, T, cout, f"{p}.convs1.{d}", rk, dil, h)
                    a2 = ops.empty((B, T, cout), self.adt, self.device)
                    ops.leaky_relu_bf16(h, 0.1, a2)
                    nxt = ops.empty((B * T, cout), F32, self.device)
                    self._conv1d(a2, B, T, cout, f"{p}.convs2.{d}", rk, 1, nxt, residual=cur)
                    cur = nxt
                if xs is None:
                    xs = cur
                else:
                    acc = ops.empty((B * T, cout), F32, self.device)
                    ops.add(xs, cur, acc)
                    xs = acc
            x = xs
            scale = 1.0 / len(HIFI_RES_K)
        a = ops.empty((B, T, x.shape[-1]), self.adt, self.device)
        ops.leaky_relu_bf16(x, 0.01, a, scale=scale)                # F.leaky_relu default slope (models.py:161)
        o = ops.empty((B * T, 1), F32, self.device)
        self._conv1d(a, B, T, x.shape[-1], "conv_post", 7, 1, o)
        wav = ops.empty((B, T), F32, self.device)
        ops.tanh(o.reshape(-1), wav)
        return wav

    def __call__(self, mel: torch.Tensor) -> torch.Tensor:
        """mel: [T, 64] or [B, T, 64] -> waveform [T_out] or [B, T_out] (SpeechT5HifiGan convention)."""
        if mel.dim() == 2:
            return self._many(mel[None])[0]
        return self._many(mel)


# ------------------------------------------------------------------------------------------------- facade
class AudioEnds:
    """VAE + vocoder of one model, built lazily (models.py wrappers call into this)."""

    def __init__(self, device, ckpt_dir: Optional[str] = None, allow_synthetic: bool = False):
        self.device = torch.device(device)
        self.ckpt_dir = ckpt_dir
        self.allow_synthetic = allow_synthetic
        self._vae: Optional[VAEEngine] = None
        self._voc: Optional[HiFiGANEngine] = None

    def _require(self, sub: str) -> bool:
        """True if <ckpt>/<sub> holds weights; otherwise seeded synthetic weights only on explicit request."""
        if self.ckpt_dir and os.path.isdir(os.path.join(self.ckpt_dir, sub)):
            return True
        if not self.allow_synthetic:
            raise FileNotFoundError(
                f"no {sub}/ weights under {self.ckpt_dir!r} and synthetic weights were not requested "
                "(allow_synthetic=True / AEDIT_ALLOW_SYNTHETIC=1)")
        return False

    def _tango_state(self):
        """TANGO snapshot (models.py:402-421): VAE + vocoder live in pytorch_model_vae.bin at the snapshot root under the
        original AudioLDM names; vae_config.json may carry the latent scale factor."""
        p = os.path.join(self.ckpt_dir, "pytorch_model_vae.bin") if self.ckpt_dir else None
        if not p or not os.path.exists(p):
            return None
        if getattr(self, "_tango_sd", None) is None:
            self._tango_sd = torch.load(p, map_location="cpu", weights_only=True)
        return self._tango_sd

    def vae(self) -> VAEEngine:
        if self._vae is None:
            w, sf = None, VAE_SCALING
            tango = self._tango_state()
            if tango is not None:
                import json
                w = ldm_vae_to_canonical(tango)
                missing = [k for k in vae_weight_shapes() if k not in w]
                if missing:
                    raise KeyError(f"pytorch_model_vae.bin lacks {len(missing)} VAE tensors, e.g. {missing[:3]}")
                cfgp = os.path.join(self.ckpt_dir, "vae_config.json")
                if os.path.exists(cfgp):
                    sf = float(json.load(open(cfgp)).get("scale_factor", 1.0))
                else:
                    sf = 1.0
            elif self._require("vae"):
                import json
                w = _load_state(os.path.join(self.ckpt_dir, "vae"))
                cfgp = os.path.join(self.ckpt_dir, "vae", "config.json")
                if os.path.exists(cfgp):
                    sf = json.load(open(cfgp)).get("scaling_factor", sf)
            self._vae = VAEEngine(self.device, w, sf)
        return self._vae

    def voc(self) -> HiFiGANEngine:
        if self._voc is None:
            w, nb = None, False
            tango = self._tango_state()
            if tango is not None and any(k.startswith("vocoder.") for k in tango):
                w = ldm_hifigan_to_canonical(tango)
            elif self._require("vocoder"):
                import json
                w = _load_state(os.path.join(self.ckpt_dir, "vocoder"))
                cfgp = os.path.join(self.ckpt_dir, "vocoder", "config.json")
                if os.path.exists(cfgp):
                    nb = bool(json.load(open(cfgp)).get("normalize_before", False))
            self._voc = HiFiGANEngine(self.device, w, normalize_before=nb)
        return self._voc

    def vae_encode_mode(self, x):
        return self.vae().encode_mode(x)

    def vae_encode_sample(self, x):
        """TANGO: posterior.sample() * scale_factor (models.py:447; distributions.py:24-73 clamps logvar to [-30, 20])."""
        mom = self.vae().encode_moments(x)
        mean, logvar = mom[:, :VAE_Z], torch.clamp(mom[:, VAE_Z:], -30.0, 20.0)
        return (mean + torch.exp(0.5 * logvar) * torch.randn_like(mean)) * self.vae().scaling

    def vae_decode(self, z):
        return self.vae().decode(z)

    def vocoder(self, mel):
        return self.voc()(mel)
