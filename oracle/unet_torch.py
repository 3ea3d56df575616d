"""TEST INFRASTRUCTURE ONLY (oracle) — fp32 CPU restatement of the conditional 2D-conv U-Net the
reference drives through `PipelineWrapper.unet_forward` (reference code/models.py:160-393 for
AudioLDM/TANGO, :691-899 for AudioLDM2).

The U-Net body itself is un-vendored third-party code (diffusers `UNet2DConditionModel` /
`AudioLDM2UNet2DConditionModel`, unpinned in requirements.txt:1).  This file restates its published
algorithm; every block is anchored on the only in-tree statement of the same math, the vendored
AudioLDM-1 U-Net:
    ResBlock            code/audioldm/latent_diffusion/openaimodel.py:175-286
    Up/Downsample       openaimodel.py:92-172
    block wiring        openaimodel.py:432-851  (skip stack: :840-846)
    SpatialTransformer  code/audioldm/latent_diffusion/attention.py:410-469
    BasicTransformerBlock attention.py:370-407, CrossAttention :149-323, GEGLU :37-44
    timestep_embedding  code/audioldm/latent_diffusion/util.py:173-197 ([cos, sin], divisor = half)
    GroupNorm32         util.py:240-242 (eps 1e-5); transformer Normalize attention.py:75-78 (eps 1e-6)
and on the top-level op order / tap points of models.py:160-393.

Parity pins: oracle/make_golden.py runs the *vendored* UNetModel (imported unmodified from
/root/reference) with weights converted by `ldm_to_canonical` and commits its outputs under
tests/golden/; tests/test_oracle_unet.py checks this restatement against those fixtures.  For the
AudioLDM2 two-stream routing and TANGO linear-projection variants no in-tree statement exists:
those branches are "parity unpinned" (restated from the published diffusers algorithm only).

Weights use the diffusers state-dict naming ("canonical" names) because that is what the
checkpoints the reference loads (models.py:478,556-564,418-422) contain.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """util.py:173-197 == diffusers Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, w, prefix, eps, groups=32):
    return F.group_norm(x, groups, w[prefix + ".weight"], w[prefix + ".bias"], eps)


def _conv(x, w, prefix, stride=1, padding=1):
    return F.conv2d(x, w[prefix + ".weight"], w[prefix + ".bias"], stride=stride, padding=padding)


def _lin(x, w, prefix, bias=True):
    return F.linear(x, w[prefix + ".weight"], w.get(prefix + ".bias") if bias else None)


def resnet_block(x, emb_act, w, p, eps, groups):
    """openaimodel.py:264-286 (use_scale_shift_norm=False): h = conv1(silu(gn(x))) + Linear(silu(emb));
    h = conv2(silu(gn(h))); return skip(x) + h.  `emb_act` is silu(emb) already."""
    h = _conv(F.silu(_gn(x, w, p + ".norm1", eps, groups)), w, p + ".conv1")
    h = h + _lin(emb_act, w, p + ".time_emb_proj")[:, :, None, None]
    h = _conv(F.silu(_gn(h, w, p + ".norm2", eps, groups)), w, p + ".conv2")
    if (p + ".conv_shortcut.weight") in w:
        x = _conv(x, w, p + ".conv_shortcut", padding=0)
    return x + h


def attention(q_in, kv_in, w, p, heads, bias=None):
    """attention.py:220-323: q/k/v Linear (no bias), scale d_head**-0.5, softmax over keys, to_out.0 Linear.
    `bias` is the additive key mask [B, 1, Lk] of models.py:204-210 (keep 0 / discard -10000)."""
    q = _lin(q_in, w, p + ".to_q", bias=False)
    k = _lin(kv_in, w, p + ".to_k", bias=False)
    v = _lin(kv_in, w, p + ".to_v", bias=False)
    B, Nq, C = q.shape
    d = C // heads
    q = q.view(B, Nq, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    if bias is not None:
        s = s + bias[:, None, :, :]
    o = torch.matmul(s.softmax(dim=-1), v)
    o = o.transpose(1, 2).reshape(B, Nq, C)
    return _lin(o, w, p + ".to_out.0")


def transformer_2d(x, w, p, heads, n_layers, linear_proj, ctx, ctx_bias, groups):
    """attention.py:455-469 (SpatialTransformer) around :402-407 (BasicTransformerBlock):
    x + proj_out(blocks(proj_in(gn(x)))); block: attn1(LN1) + x; attn2(LN2, ctx) + x; ff(LN3) + x."""
    B, C, H, W = x.shape
    h = _gn(x, w, p + ".norm", 1e-6, groups)
    if linear_proj:
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = _lin(h, w, p + ".proj_in")
    else:
        h = _conv(h, w, p + ".proj_in", padding=0)
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, -1)
    for l in range(n_layers):
        q = f"{p}.transformer_blocks.{l}"
        n1 = F.layer_norm(h, h.shape[-1:], w[q + ".norm1.weight"], w[q + ".norm1.bias"])
        h = attention(n1, n1, w, q + ".attn1", heads) + h
        n2 = F.layer_norm(h, h.shape[-1:], w[q + ".norm2.weight"], w[q + ".norm2.bias"])
        h = attention(n2, n2 if ctx is None else ctx, w, q + ".attn2", heads,
                      None if ctx is None else ctx_bias) + h
        n3 = F.layer_norm(h, h.shape[-1:], w[q + ".norm3.weight"], w[q + ".norm3.bias"])
        ff = _lin(n3, w, q + ".ff.net.0.proj")
        a, g = ff.chunk(2, dim=-1)
        h = _lin(a * F.gelu(g), w, q + ".ff.net.2") + h
    if linear_proj:
        h = _lin(h, w, p + ".proj_out")
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    else:
        h = h.reshape(B, H, W, -1).permute(0, 3, 1, 2)
        h = _conv(h, w, p + ".proj_out", padding=0)
    return h + x


def unet_forward(cfg, w: Dict[str, torch.Tensor], sample: torch.Tensor, timesteps: torch.Tensor,
                 streams: Sequence[Optional[torch.Tensor]] = (), stream_masks: Sequence[Optional[torch.Tensor]] = (),
                 class_labels: Optional[torch.Tensor] = None,
                 mid_block_additional_residual=None, replace_h_space=None, replace_skip_conns=None,
                 zero_out_resconns=None, probe=None):
    """Top-level order follows models.py:216-393 exactly (time emb → class emb concat → conv_in → down
    → mid → h-space tap/replace → +mid residual → up with skip replace/zero → GN/SiLU/conv_out).

    cfg: any object with the fields of audioeditingcode_b200.unet.UNetConfig (duck-typed so the oracle
    does not import the product).  streams[i]: [B, L_i, D_i] text stream i; stream_masks[i]: [B, L_i] 1=keep.
    probe(name, tensor NCHW): optional callback after conv_in and every block (error attribution, tools/).
    Returns (eps [B,Cout,H,W], h_space, extracted_res_conns dict)."""
    if probe is None:
        probe = lambda name, t: None
    B = sample.shape[0]
    G = cfg.norm_num_groups
    eps = cfg.norm_eps
    ch = cfg.block_out_channels
    nlev = len(ch)
    if timesteps.dim() == 0:
        timesteps = timesteps[None]
    timesteps = timesteps.expand(B)
    biases = [None if m is None else ((1 - m.to(sample.dtype)) * -10000.0)[:, None, :] for m in stream_masks]

    emb = timestep_embedding(timesteps, ch[0])
    emb = _lin(F.silu(_lin(emb, w, "time_embedding.linear_1")), w, "time_embedding.linear_2")
    if cfg.class_embed_dim is not None:
        cemb = _lin(class_labels, w, "class_embedding")
        emb = torch.cat([emb, cemb], dim=-1) if cfg.class_embeddings_concat else emb + cemb
    emb_act = F.silu(emb)

    def attn_site(x, p_attn_base, idx0, level):
        for j, spec in enumerate(cfg.transformer_specs):
            p = f"{p_attn_base}.{idx0 * len(cfg.transformer_specs) + j}"
            if spec is None:
                x = transformer_2d(x, w, p, cfg.num_heads[level], cfg.transformer_layers_per_block,
                                   cfg.use_linear_projection, None, None, G)
            else:
                _, si = spec
                x = transformer_2d(x, w, p, cfg.num_heads[level], cfg.transformer_layers_per_block,
                                   cfg.use_linear_projection, streams[si],
                                   biases[si] if si < len(biases) else None, G)
        return x

    h = _conv(sample, w, "conv_in")
    probe("conv_in", h)
    skips = [h]
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            h = resnet_block(h, emb_act, w, f"down_blocks.{i}.resnets.{j}", eps, G)
            probe(f"down_blocks.{i}.resnets.{j}", h)
            if cfg.attn_levels[i]:
                h = attn_site(h, f"down_blocks.{i}.attentions", j, i)
                probe(f"down_blocks.{i}.attentions.{j}", h)
            skips.append(h)
        if i != nlev - 1:
            h = _conv(h, w, f"down_blocks.{i}.downsamplers.0.conv", stride=2)
            probe(f"down_blocks.{i}.downsamplers.0", h)
            skips.append(h)

    h = resnet_block(h, emb_act, w, "mid_block.resnets.0", eps, G)
    probe("mid_block.resnets.0", h)
    h = attn_site(h, "mid_block.attentions", 0, nlev - 1)
    probe("mid_block.attentions.0", h)
    h = resnet_block(h, emb_act, w, "mid_block.resnets.1", eps, G)
    probe("mid_block.resnets.1", h)

    if replace_h_space is None:
        h_space = h.clone()
    else:
        h_space = replace_h_space
        h = replace_h_space.clone()
    if mid_block_additional_residual is not None:
        h = h + mid_block_additional_residual

    extracted = {}
    n_up = cfg.layers_per_block + 1
    for i in range(nlev):
        level = nlev - 1 - i
        res = skips[-n_up:]
        skips = skips[:-n_up]
        if replace_skip_conns is not None and replace_skip_conns.get(i):
            res = replace_skip_conns.get(i)
        if zero_out_resconns is not None:
            if (type(zero_out_resconns) is int and i >= (zero_out_resconns - 1)) or \
                    (type(zero_out_resconns) is list and i in zero_out_resconns):
                res = [torch.zeros_like(x) for x in res]
        extracted[i] = res
        res = list(res)
        for j in range(n_up):
            h = torch.cat([h, res.pop()], dim=1)
            h = resnet_block(h, emb_act, w, f"up_blocks.{i}.resnets.{j}", eps, G)
            probe(f"up_blocks.{i}.resnets.{j}", h)
            if cfg.attn_levels[level]:
                h = attn_site(h, f"up_blocks.{i}.attentions", j, level)
                probe(f"up_blocks.{i}.attentions.{j}", h)
        if i != nlev - 1:
            if skips and skips[-1].shape[2:] != (h.shape[2] * 2, h.shape[3] * 2):
                h = F.interpolate(h, size=skips[-1].shape[2:], mode="nearest")  # models.py:365-366
            else:
                h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, w, f"up_blocks.{i}.upsamplers.0.conv")
            probe(f"up_blocks.{i}.upsamplers.0", h)

    h = F.silu(_gn(h, w, "conv_norm_out", eps, G))
    out = _conv(h, w, "conv_out")
    return out, h_space, extracted


# ------------------------------------------------------------------------------------------------
# weight-shape inventory (canonical names) — drives synthetic weight generation and the product's
# packer.  Independent of the product on purpose.
# ------------------------------------------------------------------------------------------------
def weight_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    ch = cfg.block_out_channels
    nlev = len(ch)
    ted = 4 * ch[0]
    temb_ch = 2 * ted if (cfg.class_embed_dim is not None and cfg.class_embeddings_concat) else ted
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(p, o, i, bias=True):
        s[p + ".weight"] = (o, i)
        if bias:
            s[p + ".bias"] = (o,)

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        lin(p + ".time_emb_proj", cout, temb_ch)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    def transformer(p, c, spec):
        norm(p + ".norm", c)
        if cfg.use_linear_projection:
            lin(p + ".proj_in", c, c)
            lin(p + ".proj_out", c, c)
        else:
            conv(p + ".proj_in", c, c, 1)
            conv(p + ".proj_out", c, c, 1)
        for l in range(cfg.transformer_layers_per_block):
            q = f"{p}.transformer_blocks.{l}"
            for n in ("norm1", "norm2", "norm3"):
                norm(f"{q}.{n}", c)
            kv = c if spec is None else spec[0]
            for a, kvd in (("attn1", c), ("attn2", kv)):
                lin(f"{q}.{a}.to_q", c, c, bias=False)
                lin(f"{q}.{a}.to_k", c, kvd, bias=False)
                lin(f"{q}.{a}.to_v", c, kvd, bias=False)
                lin(f"{q}.{a}.to_out.0", c, c)
            lin(f"{q}.ff.net.0.proj", 8 * c, c)
            lin(f"{q}.ff.net.2", c, 4 * c)

    def site(pbase, idx0, c):
        for j, spec in enumerate(cfg.transformer_specs):
            transformer(f"{pbase}.{idx0 * len(cfg.transformer_specs) + j}", c, spec)

    lin("time_embedding.linear_1", ted, ch[0])
    lin("time_embedding.linear_2", ted, ted)
    if cfg.class_embed_dim is not None:
        lin("class_embedding", ted, cfg.class_embed_dim)
    conv("conv_in", ch[0], cfg.in_channels, 3)
    skip_ch = [ch[0]]
    c = ch[0]
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", c, ch[i])
            c = ch[i]
            if cfg.attn_levels[i]:
                site(f"down_blocks.{i}.attentions", j, c)
            skip_ch.append(c)
        if i != nlev - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", c, c, 3)
            skip_ch.append(c)
    resnet("mid_block.resnets.0", c, c)
    site("mid_block.attentions", 0, c)
    resnet("mid_block.resnets.1", c, c)
    for i in range(nlev):
        level = nlev - 1 - i
        for j in range(cfg.layers_per_block + 1):
            resnet(f"up_blocks.{i}.resnets.{j}", c + skip_ch.pop(), ch[level])
            c = ch[level]
            if cfg.attn_levels[level]:
                site(f"up_blocks.{i}.attentions", j, c)
        if i != nlev - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("conv_norm_out", c)
    conv("conv_out", cfg.out_channels, c, 3)
    return s


def synthetic_weights(cfg, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights: fan-in scaled normals so activations stay O(1) through
    the depth of the net; norm gains ~1, biases small.  Deterministic given (cfg, seed): the generator
    is consumed in sorted-name order so any machine reproduces the same tensors."""
    g = torch.Generator().manual_seed(seed)
    shapes = weight_shapes(cfg)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith(".weight") and len(shp) == 1:      # norm gain
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) * (0.7 / math.sqrt(fan_in))
        out[name] = t.to(dtype)
    return out


def count_params(cfg) -> int:
    n = 0
    for shp in weight_shapes(cfg).values():
        k = 1
        for d in shp:
            k *= d
        n += k
    return n


def count_flops(cfg, H: int, W: int, B: int = 1, stream_lens: Sequence[int] = ()) -> Dict[str, float]:
    """Algorithmic FLOPs of one U-Net evaluation with SURVEY.md §8d's counting rule: 2 FLOP per MAC over
    every conv (out_numel·Cin·kh·kw), every linear (out_numel·in_features) and the attention QKᵀ / AV
    products (2·B·heads·Nq·Nk·d MACs); norms, activations, softmax excluded."""
    ch = cfg.block_out_channels
    nlev = len(ch)
    ted = 4 * ch[0]
    temb_ch = 2 * ted if (cfg.class_embed_dim is not None and cfg.class_embeddings_concat) else ted
    f = {"conv": 0.0, "linear": 0.0, "attn": 0.0}

    def conv(cin, cout, k, h, w_):
        f["conv"] += 2.0 * B * h * w_ * cout * cin * k * k

    def lin(rows, i, o):
        f["linear"] += 2.0 * rows * i * o

    def resnet(cin, cout, h, w_):
        conv(cin, cout, 3, h, w_)
        lin(B, temb_ch, cout)
        conv(cout, cout, 3, h, w_)
        if cin != cout:
            conv(cin, cout, 1, h, w_)

    def site(c, h, w_):
        T = h * w_
        for spec in cfg.transformer_specs:
            if cfg.use_linear_projection:
                lin(B * T, c, c); lin(B * T, c, c)
            else:
                conv(c, c, 1, h, w_); conv(c, c, 1, h, w_)
            for _ in range(cfg.transformer_layers_per_block):
                lin(B * T, c, 3 * c); lin(B * T, c, c)          # attn1 qkv + out
                f["attn"] += 2.0 * 2.0 * B * T * T * c            # QK^T + AV
                lin(B * T, c, c); lin(B * T, c, c)                # attn2 q + out
                if spec is None:
                    lin(B * T, c, 2 * c)
                    f["attn"] += 2.0 * 2.0 * B * T * T * c
                else:
                    L = stream_lens[spec[1]]
                    lin(B * L, spec[0], 2 * c)
                    f["attn"] += 2.0 * 2.0 * B * T * L * c
                lin(B * T, c, 8 * c); lin(B * T, 4 * c, c)

    lin(B, ch[0], ted); lin(B, ted, ted)
    if cfg.class_embed_dim is not None:
        lin(B, cfg.class_embed_dim, ted)
    h, w_ = H, W
    conv(cfg.in_channels, ch[0], 3, h, w_)
    skip = [(ch[0])]
    c = ch[0]
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            resnet(c, ch[i], h, w_); c = ch[i]
            if cfg.attn_levels[i]:
                site(c, h, w_)
            skip.append(c)
        if i != nlev - 1:
            h, w_ = (h + 1) // 2, (w_ + 1) // 2
            conv(c, c, 3, h, w_)
            skip.append(c)
    resnet(c, c, h, w_); site(c, h, w_); resnet(c, c, h, w_)
    sizes = [(H, W)]
    for i in range(nlev - 1):
        sizes.append(((sizes[-1][0] + 1) // 2, (sizes[-1][1] + 1) // 2))
    for i in range(nlev):
        level = nlev - 1 - i
        h, w_ = sizes[level]
        for j in range(cfg.layers_per_block + 1):
            resnet(c + skip.pop(), ch[level], h, w_); c = ch[level]
            if cfg.attn_levels[level]:
                site(c, h, w_)
        if i != nlev - 1:
            h2, w2 = sizes[level - 1]
            conv(c, c, 3, h2, w2)
    conv(c, cfg.out_channels, 3, H, W)
    f["total"] = f["conv"] + f["linear"] + f["attn"]
    return f


# ------------------------------------------------------------------------------------------------
# name conversion: vendored LDM UNetModel state_dict  <->  canonical (diffusers) names
# ------------------------------------------------------------------------------------------------
def ldm_name_map(cfg) -> Dict[str, str]:
    """canonical prefix -> vendored-UNetModel prefix (openaimodel.py:571-783 module order)."""
    m: Dict[str, str] = {}
    ch = cfg.block_out_channels
    nlev = len(ch)
    m["time_embedding.linear_1"] = "time_embed.0"
    m["time_embedding.linear_2"] = "time_embed.2"
    if cfg.class_embed_dim is not None:
        m["class_embedding"] = "film_emb"
    m["conv_in"] = "input_blocks.0.0"

    def res(cp, lp):
        m[cp + ".norm1"] = lp + ".in_layers.0"
        m[cp + ".conv1"] = lp + ".in_layers.2"
        m[cp + ".time_emb_proj"] = lp + ".emb_layers.1"
        m[cp + ".norm2"] = lp + ".out_layers.0"
        m[cp + ".conv2"] = lp + ".out_layers.3"
        m[cp + ".conv_shortcut"] = lp + ".skip_connection"

    n = 1
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            res(f"down_blocks.{i}.resnets.{j}", f"input_blocks.{n}.0")
            if cfg.attn_levels[i]:
                m[f"down_blocks.{i}.attentions.{j}"] = f"input_blocks.{n}.1"
            n += 1
        if i != nlev - 1:
            m[f"down_blocks.{i}.downsamplers.0.conv"] = f"input_blocks.{n}.0.op"
            n += 1
    res("mid_block.resnets.0", "middle_block.0")
    m["mid_block.attentions.0"] = "middle_block.1"
    res("mid_block.resnets.1", "middle_block.2")
    n = 0
    for i in range(nlev):
        level = nlev - 1 - i
        for j in range(cfg.layers_per_block + 1):
            res(f"up_blocks.{i}.resnets.{j}", f"output_blocks.{n}.0")
            k = 1
            if cfg.attn_levels[level]:
                m[f"up_blocks.{i}.attentions.{j}"] = f"output_blocks.{n}.1"
                k = 2
            if i != nlev - 1 and j == cfg.layers_per_block:
                m[f"up_blocks.{i}.upsamplers.0.conv"] = f"output_blocks.{n}.{k}.conv"
            n += 1
    m["conv_norm_out"] = "out.0"
    m["conv_out"] = "out.2"
    return m


def canonical_to_ldm(cfg, w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Rename canonical weights into the vendored UNetModel's state_dict naming (single transformer
    per site only — the vendored model has exactly one SpatialTransformer per attention site)."""
    assert len(cfg.transformer_specs) == 1
    pm = sorted(ldm_name_map(cfg).items(), key=lambda kv: -len(kv[0]))
    out = {}
    for name, t in w.items():
        for cp, lp in pm:
            if name == cp or name.startswith(cp + "."):
                out[lp + name[len(cp):]] = t
                break
        else:
            raise KeyError(name)
    return out
