"""Algorithmic FLOP accounting of one U-Net evaluation (SURVEY.md §8d counting rule) — used by bench.py for the
roofline numbers.  2 FLOP per MAC over every conv (out_numel*Cin*kh*kw), every linear (out_numel*in_features) and the
attention QK^T / AV products; norms, activations, softmax excluded.  Verified against SURVEY.md Appendix A.1
(AudioLDM-S @ [1,8,256,16]: conv 64.2 + linear 27.3 + attn 11.9 = 103.4 GFLOP) in tests/test_host_logic.py."""
from __future__ import annotations

from typing import Dict, Sequence


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
