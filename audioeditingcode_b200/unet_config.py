"""Architecture description of the conditional 2D-conv U-Net the hot loop drives
(reference: code/models.py:160-393 / :691-899 call into `self.model.unet.*`; hyper-parameters come from
the checkpoint's unet/config.json — [UPSTREAM] diffusers fields).

One dataclass covers the three families the reference supports on this path:
  AudioLDM   class-embedding (CLAP, 512-d) concatenated to the time embedding, self-attention-only
             transformers (vendored twin: code/audioldm/utils.py:142-157 + openaimodel.py:432-851)
  AudioLDM2  three transformers per attention site: self-only, cross->stream0 (GPT-2, 768-d),
             cross->stream1 (T5, 1024-d, masked)   (models.py:706-710,812-821)
  TANGO      SD-2.1 widths, linear projections, one cross-attention transformer per site (T5 1024-d)
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, asdict
from typing import Optional, Tuple

Spec = Optional[Tuple[int, int]]  # None = attn2 is self-attention; (cross_dim, stream_index) otherwise


@dataclass(frozen=True)
class UNetConfig:
    name: str = "custom"
    in_channels: int = 8
    out_channels: int = 8
    block_out_channels: Tuple[int, ...] = (128, 256, 384, 640)
    layers_per_block: int = 2
    attn_levels: Tuple[bool, ...] = (False, True, True, True)
    num_heads: Tuple[int, ...] = (4, 8, 12, 20)          # attention heads per level
    transformer_specs: Tuple[Spec, ...] = (None,)
    transformer_layers_per_block: int = 1
    use_linear_projection: bool = False
    class_embed_dim: Optional[int] = None               # "simple_projection" input dim
    class_embeddings_concat: bool = False
    norm_eps: float = 1e-5
    norm_num_groups: int = 32
    prediction_type: str = "epsilon"                    # scheduler-side, kept here for presets
    beta_start: float = 0.0015
    beta_end: float = 0.0195

    @property
    def n_streams(self) -> int:
        return 1 + max([s[1] for s in self.transformer_specs if s is not None], default=-1)

    def to_json(self) -> str:
        return json.dumps(asdict(self))


def audioldm(width: int = 128, head_dim: Optional[int] = None, name: str = "audioldm-s") -> UNetConfig:
    """AudioLDM-1 (audioldm/utils.py:142-157: channel_mult [1,2,3,5], 2 res blocks, attention at ds 2/4/8, FiLM concat
    of a 512-d CLAP vector).  Heads: the reference's hot path runs the diffusers UNet2DConditionModel of the converted
    checkpoints, whose config carries `attention_head_dim: 8` = EIGHT HEADS at every level ([UPSTREAM], diffusers
    reads that field as the head count) — not the original LDM's num_head_channels 32 (heads 4/8/12/20), which
    `head_dim=32` still selects (the vendored UNetModel golden uses it).  Weight shapes are the same either way; real
    checkpoints always take the split from their own unet/config.json (models.py requires it)."""
    ch = tuple(width * m for m in (1, 2, 3, 5))
    heads = (8,) * len(ch) if head_dim is None else tuple(c // head_dim for c in ch)
    return UNetConfig(name=name, block_out_channels=ch, num_heads=heads,
                      transformer_specs=(None,), class_embed_dim=512, class_embeddings_concat=True)


def audioldm2(width: int = 128, name: str = "audioldm2") -> UNetConfig:
    """AudioLDM2 [UPSTREAM]: 8 heads at every level, cross_attention_dim [None, 768, 1024] per site."""
    ch = tuple(width * m for m in (1, 2, 3, 5))
    return UNetConfig(name=name, block_out_channels=ch, num_heads=(8, 8, 8, 8),
                      transformer_specs=(None, (768, 0), (1024, 1)))


def tango(name: str = "tango") -> UNetConfig:
    """TANGO [UPSTREAM tango/configs/diffusion_model_config.json]: SD-2.1 widths, attention on levels 0-2,
    linear projections, v-prediction scheduler with SD betas (models.py:431-434)."""
    return UNetConfig(name=name, block_out_channels=(320, 640, 1280, 1280), attn_levels=(True, True, True, False),
                      num_heads=(5, 10, 20, 20), transformer_specs=((1024, 0),), use_linear_projection=True,
                      prediction_type="v_prediction", beta_start=0.00085, beta_end=0.012)


PRESETS = {
    "audioldm-s": lambda: audioldm(128, None, "audioldm-s"),
    "audioldm-m": lambda: audioldm(192, None, "audioldm-m"),
    "audioldm-l": lambda: audioldm(256, 64, "audioldm-l"),
    "audioldm2": lambda: audioldm2(128, "audioldm2"),
    # width of -large is UNVERIFIED upstream; 192 reproduces the ~750 M U-Net
    # parameter count on the model card, 256 would give ~1.4 B.
    "audioldm2-large": lambda: audioldm2(192, "audioldm2-large"),
    "tango": lambda: tango(),
    # small nets for tests (channels kept multiples of 64 so the TMA implicit-conv path is exercised)
    "tiny-audioldm": lambda: UNetConfig(name="tiny-audioldm", block_out_channels=(64, 128), layers_per_block=1,
                                        attn_levels=(False, True), num_heads=(2, 4), class_embed_dim=512,
                                        class_embeddings_concat=True),
    "tiny-audioldm2": lambda: UNetConfig(name="tiny-audioldm2", block_out_channels=(64, 128), layers_per_block=1,
                                         attn_levels=(False, True), num_heads=(2, 4),
                                         transformer_specs=(None, (96, 0), (160, 1))),
    "tiny-tango": lambda: UNetConfig(name="tiny-tango", block_out_channels=(64, 128), layers_per_block=1,
                                     attn_levels=(True, False), num_heads=(2, 4), transformer_specs=((160, 0),),
                                     use_linear_projection=True, prediction_type="v_prediction",
                                     beta_start=0.00085, beta_end=0.012),
}


def preset(name: str) -> UNetConfig:
    return PRESETS[name]()


def from_model_id(model_id: str) -> UNetConfig:
    """Same substring dispatch as load_model (models.py:1357-1364)."""
    mid = model_id.lower()
    if "tango" in mid:
        return preset("tiny-tango") if "tiny" in mid else preset("tango")
    if "audioldm2" in mid:
        if "tiny" in mid:
            return preset("tiny-audioldm2")
        return preset("audioldm2-large") if "large" in mid else preset("audioldm2")
    if "audioldm" in mid:
        if "tiny" in mid:
            return preset("tiny-audioldm")
        if "-l-" in mid or mid.endswith("-l"):
            return preset("audioldm-l")
        if "-m-" in mid or mid.endswith("-m"):
            return preset("audioldm-m")
        return preset("audioldm-s")
    raise ValueError(f"unsupported model id for the audio hot path: {model_id}")
