"""Host-logic test (CPU): UNetEngine's wiring — weight packing, fused projections, skip bookkeeping, text cache,
taps — checked against the oracle restatement with an fp32 torch stand-in for the device ops (tests/torch_ops.py).
The CUDA kernels themselves are checked on the GPU box (tests/test_gpu_*.py)."""
import pytest
import torch

from oracle import unet_torch as U
from audioeditingcode_b200 import unet_config as C
from audioeditingcode_b200.unet import UNetEngine
from tests.torch_ops import TorchOps


def _streams(cfg, R, lens=(8, 5)):
    g = torch.Generator().manual_seed(4)
    streams, masks = [], []
    dims = {}
    for s in cfg.transformer_specs:
        if s is not None:
            dims[s[1]] = s[0]
    for i in range(cfg.n_streams):
        L = lens[i % len(lens)]
        streams.append(torch.randn(R, L, dims[i], generator=g))
        m = torch.ones(R, L)
        if i == cfg.n_streams - 1 and L > 2:
            m[0, L - 2:] = 0           # right-padded mask on one row
        masks.append(m)
    return streams, masks


@pytest.mark.parametrize("name,H,W", [("tiny-audioldm", 16, 16), ("tiny-audioldm2", 16, 16), ("tiny-tango", 16, 16),
                                      ("tiny-audioldm", 20, 16)])
def test_engine_matches_oracle(name, H, W):
    cfg = C.preset(name)
    w = U.synthetic_weights(cfg, seed=0)
    eng = UNetEngine(cfg, w, "cpu", ops=TorchOps())
    g = torch.Generator().manual_seed(3)
    B = 3
    x = torch.randn(B, 8, H, W, generator=g)
    t = torch.tensor([981, 441, 1])
    kw_o, kw_e = {}, {}
    if cfg.class_embed_dim is not None:
        y = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1)
        kw_o["class_labels"] = y
        kw_e["class_labels"] = y
    if cfg.n_streams:
        R = 2
        streams, masks = _streams(cfg, R)
        slot = torch.tensor([0, 1, 1], dtype=torch.int32)
        kw_o["streams"] = [s[slot.long()] for s in streams]
        kw_o["stream_masks"] = [m[slot.long()] for m in masks]
        kw_e["text"] = eng.prepare_text(streams, masks)
        kw_e["slot_map"] = slot
    with torch.no_grad():
        ref, hs_ref, ex_ref = U.unet_forward(cfg, w, x, t, **kw_o)
        out, hs, ex = eng.forward(x, t, want_taps=True, **kw_e)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4), (out - ref).abs().max()
    assert torch.allclose(hs, hs_ref, atol=2e-5, rtol=1e-4)
    for i in ex_ref:
        for a, b in zip(ex[i], ex_ref[i]):
            assert torch.allclose(a, b, atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("name,lens", [("tiny-audioldm2", (8, 5)), ("tiny-tango", (13,)), ("tiny-audioldm2", (8, 16))])
def test_folded_cross_attention_wiring_matches_oracle(name, lens):
    """The algebra and index layout of the folded cross-attention (UNetEngine._fold_cross_attention: KW = scale K.Wq per
    (text row, head), VW = Wo.V^T, grouped softmax, per-sample text row selection, key mask / padding bias) checked in
    fp32 on the CPU against the oracle's ordinary attention — independent of the CUDA kernels."""
    cfg = C.preset(name)
    w = U.synthetic_weights(cfg, seed=0)
    eng = UNetEngine(cfg, w, "cpu", ops=TorchOps())
    eng.fold_cross_attn = True
    g = torch.Generator().manual_seed(3)
    B, H, W = 3, 16, 16
    x = torch.randn(B, 8, H, W, generator=g)
    t = torch.tensor([981, 441, 1])
    streams, masks = _streams(cfg, 2, lens)
    slot = torch.tensor([0, 1, 1], dtype=torch.int32)
    text = eng.prepare_text(streams, masks)
    assert text.folded, "no cross-attention layer was folded"
    with torch.no_grad():
        ref = U.unet_forward(cfg, w, x, t, streams=[s[slot.long()] for s in streams],
                             stream_masks=[m[slot.long()] for m in masks])[0]
        out = eng.forward(x, t, text=text, slot_map=slot)
    assert torch.allclose(out, ref, atol=5e-5, rtol=2e-4), (out - ref).abs().max()


@pytest.mark.parametrize("name,H", [("tiny-audioldm", 32), ("tiny-audioldm2", 32), ("tiny-audioldm", 20)])
def test_groupnorm_column_statistics_bookkeeping(name, H):
    """GroupNorm statistics handed over by the producing GEMMs (UNetEngine._cs_begin / _cs_take / _cs_of: one zeroed
    arena per evaluation, a slice per GEMM output, skip connections consumed much later, concatenated inputs, tap paths
    that must NOT find statistics) — the bookkeeping checked on the CPU against the oracle, taps included."""
    cfg = C.preset(name)
    w = U.synthetic_weights(cfg, seed=0)
    ops = TorchOps()
    eng = UNetEngine(cfg, w, "cpu", ops=ops)
    eng.gn_colstats = True
    seen = {"cs": 0, "plain": 0}
    orig = ops.groupnorm

    def spy(*a, cs1=None, cs2=None, **k):
        seen["cs" if cs1 is not None else "plain"] += 1
        return orig(*a, cs1=cs1, cs2=cs2, **k)
    ops.groupnorm = spy
    g = torch.Generator().manual_seed(3)
    B = 2
    x = torch.randn(B, 8, H, 16, generator=g)
    t = torch.tensor([981, 1])
    kw_o, kw_e = {}, {}
    if cfg.class_embed_dim is not None:
        y = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1)
        kw_o["class_labels"] = kw_e["class_labels"] = y
    if cfg.n_streams:
        streams, masks = _streams(cfg, B)
        kw_o["streams"], kw_o["stream_masks"] = streams, masks
        kw_e["text"] = eng.prepare_text(streams, masks)
        kw_e["slot_map"] = torch.arange(B, dtype=torch.int32)
    with torch.no_grad():
        ref, hs_ref, _ = U.unet_forward(cfg, w, x, t, **kw_o)
        out, hs, _ = eng.forward(x, t, want_taps=True, **kw_e)
    assert torch.allclose(out, ref, atol=1e-4, rtol=5e-4), (out - ref).abs().max()
    assert torch.allclose(hs, hs_ref, atol=1e-4, rtol=5e-4)
    if H % 32 == 0 or (H * 16) % 32 == 0:
        assert seen["cs"] > 0, "no GroupNorm used the producers' statistics"
    # tap paths: replaced / added / zeroed tensors carry no statistics and must take the statistics pass
    with torch.no_grad():
        add = 0.1 * torch.randn(hs_ref.shape, generator=g)
        ref2 = U.unet_forward(cfg, w, x, t, mid_block_additional_residual=add, zero_out_resconns=[0], **kw_o)[0]
        out2 = eng.forward(x, t, mid_block_additional_residual=add, zero_out_resconns=[0], **kw_e)
    assert torch.allclose(out2, ref2, atol=1e-4, rtol=5e-4), (out2 - ref2).abs().max()


def test_engine_taps_inject():
    cfg = C.preset("tiny-audioldm")
    w = U.synthetic_weights(cfg, seed=0)
    eng = UNetEngine(cfg, w, "cpu", ops=TorchOps())
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 8, 16, 16, generator=g)
    t = torch.tensor([501])
    y = torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1)
    with torch.no_grad():
        _, hs, ex = U.unet_forward(cfg, w, x, t, class_labels=y)
        add = 0.1 * torch.randn(hs.shape, generator=g)
        rep = torch.randn(hs.shape, generator=g)
        skip_rep = {1: [torch.randn(s.shape, generator=g) for s in ex[1]]}
        ref = U.unet_forward(cfg, w, x, t, class_labels=y, mid_block_additional_residual=add, replace_h_space=rep,
                             replace_skip_conns=skip_rep, zero_out_resconns=[0])[0]
        out = eng.forward(x, t, class_labels=y, mid_block_additional_residual=add, replace_h_space=rep,
                          replace_skip_conns=skip_rep, zero_out_resconns=[0])
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4)
