"""Weight inventory, synthetic initialisation and checkpoint loading for the U-Net engine.

Checkpoints the reference loads (code/models.py:478,556-564 `from_pretrained`; :414-422 TANGO `.bin`) store the
U-Net under diffusers state-dict names; those names are the canonical keys everywhere in this package.
There is no network in the build/bench environment, so `synthetic_weights` provides seeded stand-ins of the
exact architecture (the benchmark contract is random-init weights of the named architecture).
"""
from __future__ import annotations

import json
import math
import os
from typing import Optional, Dict, Tuple

import torch

from .unet_config import UNetConfig


def weight_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    ch = cfg.block_out_channels
    nlev = len(ch)
    ted = 4 * ch[0]
    temb_ch = 2 * ted if (cfg.class_embed_dim is not None and cfg.class_embeddings_concat) else ted
    shapes: Dict[str, Tuple[int, ...]] = {}

    def lin(p, o, i, bias=True):
        shapes[p + ".weight"] = (o, i)
        if bias:
            shapes[p + ".bias"] = (o,)

    def conv(p, o, i, k):
        shapes[p + ".weight"] = (o, i, k, k)
        shapes[p + ".bias"] = (o,)

    def norm(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)

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
            kvd = c if spec is None else spec[0]
            for a, d in (("attn1", c), ("attn2", kvd)):
                lin(f"{q}.{a}.to_q", c, c, bias=False)
                lin(f"{q}.{a}.to_k", c, d, bias=False)
                lin(f"{q}.{a}.to_v", c, d, bias=False)
                lin(f"{q}.{a}.to_out.0", c, c)
            lin(f"{q}.ff.net.0.proj", 8 * c, c)
            lin(f"{q}.ff.net.2", c, 4 * c)

    def site(base, idx0, c):
        ns = len(cfg.transformer_specs)
        for j, spec in enumerate(cfg.transformer_specs):
            transformer(f"{base}.{idx0 * ns + j}", c, spec)

    lin("time_embedding.linear_1", ted, ch[0])
    lin("time_embedding.linear_2", ted, ted)
    if cfg.class_embed_dim is not None:
        lin("class_embedding", ted, cfg.class_embed_dim)
    conv("conv_in", ch[0], cfg.in_channels, 3)
    skip = [ch[0]]
    c = ch[0]
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", c, ch[i])
            c = ch[i]
            if cfg.attn_levels[i]:
                site(f"down_blocks.{i}.attentions", j, c)
            skip.append(c)
        if i != nlev - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", c, c, 3)
            skip.append(c)
    resnet("mid_block.resnets.0", c, c)
    site("mid_block.attentions", 0, c)
    resnet("mid_block.resnets.1", c, c)
    for i in range(nlev):
        level = nlev - 1 - i
        for j in range(cfg.layers_per_block + 1):
            resnet(f"up_blocks.{i}.resnets.{j}", c + skip.pop(), ch[level])
            c = ch[level]
            if cfg.attn_levels[level]:
                site(f"up_blocks.{i}.attentions", j, c)
        if i != nlev - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("conv_norm_out", c)
    conv("conv_out", cfg.out_channels, c, 3)
    return shapes


def synthetic_weights(cfg: UNetConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded fp32 CPU weights: fan-in-scaled normals (activations stay O(1) through the depth), norm gains
    around 1, small biases.  The generator is consumed in sorted-name order so every machine reproduces the
    same tensors (tests compare against oracle.unet_torch.synthetic_weights)."""
    g = torch.Generator().manual_seed(seed)
    shapes = weight_shapes(cfg)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith(".weight") and len(shp) == 1:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) * (0.7 / math.sqrt(fan_in))
        out[name] = t
    return out


def count_params(cfg: UNetConfig) -> int:
    return sum(math.prod(s) for s in weight_shapes(cfg).values())


def load_unet_checkpoint(path: str, cfg: UNetConfig) -> Dict[str, torch.Tensor]:
    """Load a diffusers-format U-Net state dict from `path` (a directory with
    diffusion_pytorch_model.safetensors / .bin, or a single file; TANGO's pytorch_model_main.bin carries the
    same names under a `unet.` prefix, models.py:418-422)."""
    files = [path] if os.path.isfile(path) else [
        os.path.join(path, f) for f in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin",
                                        "pytorch_model_main.bin") if os.path.exists(os.path.join(path, f))]
    if not files:
        raise FileNotFoundError(f"no U-Net checkpoint under {path}")
    f = files[0]
    if f.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(f)
    else:
        sd = torch.load(f, map_location="cpu", weights_only=True)
    if any(k.startswith("unet.") for k in sd):
        sd = {k[len("unet."):]: v for k, v in sd.items() if k.startswith("unet.")}
    want = weight_shapes(cfg)
    missing = [k for k in want if k not in sd]
    if missing:
        raise KeyError(f"checkpoint {f} lacks {len(missing)} tensors, e.g. {missing[:3]}")
    out = {}
    for k, shp in want.items():
        t = sd[k].float()
        if tuple(t.shape) != tuple(shp):
            raise ValueError(f"{k}: checkpoint shape {tuple(t.shape)} != architecture {shp}")
        out[k] = t
    return out


def unet_config_from_json(path: str, name: str = "checkpoint", scheduler_json: Optional[str] = None) -> UNetConfig:
    """[UPSTREAM] diffusers unet/config.json -> UNetConfig (fields per SURVEY.md Appendix B).  diffusers treats
    `attention_head_dim` as the NUMBER of heads when `num_attention_heads` is absent (UNet2DConditionModel.__init__);
    so does this.  scheduler_json (scheduler/scheduler_config.json): beta range and prediction type."""
    with open(path) as fh:
        j = json.load(fh)
    sched = {}
    if scheduler_json and os.path.exists(scheduler_json):
        with open(scheduler_json) as fh:
            sj = json.load(fh)
        sched = {k: sj[k] for k in ("beta_start", "beta_end", "prediction_type") if k in sj}
    ch = tuple(j["block_out_channels"])
    down = j["down_block_types"]
    attn = tuple("CrossAttn" in d for d in down)
    heads = j.get("num_attention_heads") or j.get("attention_head_dim")
    heads = tuple(heads) if isinstance(heads, (list, tuple)) else (heads,) * len(ch)
    cad = j.get("cross_attention_dim")
    if isinstance(cad, (list, tuple)) and isinstance(cad[0], (list, tuple)):     # AudioLDM2 nested list
        specs, si = [], 0
        for d in cad[0]:
            if d is None:
                specs.append(None)
            else:
                specs.append((int(d), si))
                si += 1
        specs = tuple(specs)
    elif j.get("class_embed_type") == "simple_projection":                        # AudioLDM: attn2 = self
        specs = (None,)
    else:
        specs = ((int(cad if not isinstance(cad, (list, tuple)) else cad[0]), 0),)
    return UNetConfig(name=name, in_channels=j["in_channels"], out_channels=j["out_channels"], block_out_channels=ch,
                      layers_per_block=j.get("layers_per_block", 2), attn_levels=attn, num_heads=heads,
                      transformer_specs=specs, transformer_layers_per_block=j.get("transformer_layers_per_block", 1),
                      use_linear_projection=bool(j.get("use_linear_projection", False)),
                      class_embed_dim=j.get("projection_class_embeddings_input_dim")
                      if j.get("class_embed_type") == "simple_projection" else None,
                      class_embeddings_concat=bool(j.get("class_embeddings_concat", False)),
                      norm_eps=j.get("norm_eps", 1e-5), norm_num_groups=j.get("norm_num_groups", 32), **sched)
