"""TEST INFRASTRUCTURE ONLY (oracle) — fp32 CPU restatements of the audio "ends" of the path
(SURVEY.md §8 rows a10, a11):

  VAE   [UPSTREAM] diffusers AutoencoderKL as used by models.py:495-503 (`encode(x).latent_dist.mode()*scaling_factor`,
        `decode(z/scaling_factor).sample`); in-tree statement of the same network:
        code/audioldm/variational_autoencoder/modules.py:419-543 (Encoder), :546-683 (Decoder), :118-175 (ResnetBlock),
        :185-230 (AttnBlock), :76-94 (Downsample: pad (0,1,0,1) + stride-2 conv), autoencoder.py:35-36,49-61
        (quant_conv / post_quant_conv); GroupNorm(32, eps=1e-6).
  HiFi-GAN  [UPSTREAM] transformers SpeechT5HifiGan (verified against the installed transformers source);
        in-tree statement: code/audioldm/hifigan/models.py:20-165 with the 16 kHz / 64-mel config of
        hifigan/utilities.py:9-39.

Weights use the diffusers / transformers state-dict names.  `*_to_ldm` rename them into the vendored modules'
naming so oracle/make_golden.py can run the UNMODIFIED vendored modules on the same weights; the restatements here
are pinned against those golden outputs in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

VAE_CH = 128
VAE_MULT = (1, 2, 4)
VAE_RES = 2
VAE_Z = 8
VAE_SCALING = 0.9227914214134216   # [UPSTREAM] cvssp/audioldm* vae/config.json scaling_factor

HIFI_RATES = (5, 4, 2, 2, 2)
HIFI_KERNELS = (16, 16, 8, 4, 4)
HIFI_INIT = 1024
HIFI_RES_K = (3, 7, 11)
HIFI_RES_D = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
HIFI_MELS = 64


# ------------------------------------------------------------------------------------------------- VAE
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
        norm(p + ".norm1", ci); conv(p + ".conv1", co, ci, 3); norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
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


def _synth(shapes, seed):
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


def vae_synthetic_weights(seed: int = 0):
    return _synth(vae_weight_shapes(), seed)


def _gn(x, w, p):
    return F.group_norm(x, 32, w[p + ".weight"], w[p + ".bias"], 1e-6)


def _conv(x, w, p, stride=1, padding=1):
    return F.conv2d(x, w[p + ".weight"], w[p + ".bias"], stride=stride, padding=padding)


def _res(x, w, p):                                      # modules.py:152-175 (temb is None)
    h = _conv(F.silu(_gn(x, w, p + ".norm1")), w, p + ".conv1")
    h = _conv(F.silu(_gn(h, w, p + ".norm2")), w, p + ".conv2")
    if (p + ".conv_shortcut.weight") in w:
        x = _conv(x, w, p + ".conv_shortcut", padding=0)
    return x + h


def _attn(x, w, p):                                     # modules.py:203-230: single head over H*W tokens
    B, C, H, W = x.shape
    h = _gn(x, w, p + ".group_norm").permute(0, 2, 3, 1).reshape(B, H * W, C)
    q = F.linear(h, w[p + ".to_q.weight"], w[p + ".to_q.bias"])
    k = F.linear(h, w[p + ".to_k.weight"], w[p + ".to_k.bias"])
    v = F.linear(h, w[p + ".to_v.weight"], w[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) * (int(C) ** -0.5), dim=2) @ v
    a = F.linear(a, w[p + ".to_out.0.weight"], w[p + ".to_out.0.bias"])
    return x + a.reshape(B, H, W, C).permute(0, 3, 1, 2)


def _mid(x, w, p):
    x = _res(x, w, p + ".resnets.0")
    x = _attn(x, w, p + ".attentions.0")
    return _res(x, w, p + ".resnets.1")


def vae_encode_moments(w, x):
    """x: [B,1,T,64] log-mel -> moments [B,16,T/4,16] (mean = first 8 channels).  modules.py:516-543 + autoencoder.py:49-55."""
    h = _conv(x, w, "encoder.conv_in")
    n = len(VAE_MULT)
    for i in range(n):
        for j in range(VAE_RES):
            h = _res(h, w, f"encoder.down_blocks.{i}.resnets.{j}")
        if i != n - 1:
            h = F.pad(h, (0, 1, 0, 1))                  # modules.py:87-89
            h = _conv(h, w, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2, padding=0)
    h = _mid(h, w, "encoder.mid_block")
    h = _conv(F.silu(_gn(h, w, "encoder.conv_norm_out")), w, "encoder.conv_out")
    return _conv(h, w, "quant_conv", padding=0)


def vae_encode_mode(w, x, scaling=VAE_SCALING):
    return vae_encode_moments(w, x)[:, :VAE_Z] * scaling          # models.py:499


def vae_decode(w, z, scaling=VAE_SCALING):
    """models.py:503 + modules.py:654-683."""
    h = _conv(z * (1 / scaling), w, "post_quant_conv", padding=0)
    h = _conv(h, w, "decoder.conv_in")
    h = _mid(h, w, "decoder.mid_block")
    n = len(VAE_MULT)
    for i in range(n):
        for j in range(VAE_RES + 1):
            h = _res(h, w, f"decoder.up_blocks.{i}.resnets.{j}")
        if i != n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, w, f"decoder.up_blocks.{i}.upsamplers.0.conv")
    return _conv(F.silu(_gn(h, w, "decoder.conv_norm_out")), w, "decoder.conv_out")


def vae_to_ldm(w):
    """canonical (diffusers) -> vendored AutoencoderKL state_dict names."""
    out = {}
    n = len(VAE_MULT)

    def put(src, dst, conv1x1=False):
        for suf in (".weight", ".bias"):
            if src + suf in w:
                t = w[src + suf]
                if conv1x1 and suf == ".weight" and t.dim() == 2:
                    t = t[:, :, None, None]
                out[dst + suf] = t

    def res(src, dst):
        for a, b in (("norm1", "norm1"), ("conv1", "conv1"), ("norm2", "norm2"), ("conv2", "conv2"),
                     ("conv_shortcut", "nin_shortcut")):
            put(f"{src}.{a}", f"{dst}.{b}")

    def mid(src, dst):
        res(src + ".resnets.0", dst + ".block_1")
        res(src + ".resnets.1", dst + ".block_2")
        a = src + ".attentions.0"
        put(a + ".group_norm", dst + ".attn_1.norm")
        for s_, d_ in (("to_q", "q"), ("to_k", "k"), ("to_v", "v"), ("to_out.0", "proj_out")):
            put(f"{a}.{s_}", f"{dst}.attn_1.{d_}", conv1x1=True)

    put("encoder.conv_in", "encoder.conv_in")
    for i in range(n):
        for j in range(VAE_RES):
            res(f"encoder.down_blocks.{i}.resnets.{j}", f"encoder.down.{i}.block.{j}")
        put(f"encoder.down_blocks.{i}.downsamplers.0.conv", f"encoder.down.{i}.downsample.conv")
    mid("encoder.mid_block", "encoder.mid")
    put("encoder.conv_norm_out", "encoder.norm_out")
    put("encoder.conv_out", "encoder.conv_out")
    put("quant_conv", "quant_conv")
    put("post_quant_conv", "post_quant_conv")
    put("decoder.conv_in", "decoder.conv_in")
    mid("decoder.mid_block", "decoder.mid")
    for i in range(n):
        lvl = n - 1 - i
        for j in range(VAE_RES + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", f"decoder.up.{lvl}.block.{j}")
        put(f"decoder.up_blocks.{i}.upsamplers.0.conv", f"decoder.up.{lvl}.upsample.conv")
    put("decoder.conv_norm_out", "decoder.norm_out")
    put("decoder.conv_out", "decoder.conv_out")
    return out


# ------------------------------------------------------------------------------------------------- HiFi-GAN
def hifigan_weight_shapes() -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    s["conv_pre.weight"] = (HIFI_INIT, HIFI_MELS, 7)
    s["conv_pre.bias"] = (HIFI_INIT,)
    ch = HIFI_INIT
    for i, (u, k) in enumerate(zip(HIFI_RATES, HIFI_KERNELS)):
        s[f"upsampler.{i}.weight"] = (ch, ch // 2, k)          # ConvTranspose1d: [Cin, Cout, k]
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


def hifigan_synthetic_weights(seed: int = 0):
    return _synth(hifigan_weight_shapes(), seed)


def hifigan_forward(w, mel):
    """mel: [B, T, 64] log-mel (SpeechT5HifiGan input convention) -> waveform [B, T*160 (+ tail)].
    hifigan/models.py:147-165: conv_pre; per stage leaky(0.1) -> ConvTranspose1d -> mean of 3 MRF ResBlocks;
    leaky(default 0.01) -> conv_post -> tanh."""
    x = mel.transpose(1, 2)
    x = F.conv1d(x, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(HIFI_RATES, HIFI_KERNELS)):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, w[f"upsampler.{i}.weight"], w[f"upsampler.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(HIFI_RES_K, HIFI_RES_D)):
            p = f"resblocks.{i * 3 + j}"
            y = x
            for d, dil in enumerate(rd):                       # hifigan/models.py:96-103
                xt = F.leaky_relu(y, 0.1)
                xt = F.conv1d(xt, w[f"{p}.convs1.{d}.weight"], w[f"{p}.convs1.{d}.bias"], dilation=dil,
                              padding=(rk * dil - dil) // 2)
                xt = F.leaky_relu(xt, 0.1)
                xt = F.conv1d(xt, w[f"{p}.convs2.{d}.weight"], w[f"{p}.convs2.{d}.bias"], padding=(rk - 1) // 2)
                y = xt + y
            xs = y if xs is None else xs + y
        x = xs / len(HIFI_RES_K)
    x = F.leaky_relu(x)
    x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
    return torch.tanh(x).squeeze(1)


def hifigan_to_ldm(w):
    """canonical (transformers SpeechT5HifiGan) -> vendored Generator names (after remove_weight_norm)."""
    out = {}
    for k, v in w.items():
        out[k.replace("upsampler.", "ups.")] = v
    return out
