"""GPU parity tests of the individual libaedit kernels (run with `-m gpu` on a B200).  Each kernel is compared
with a plain PyTorch fp32 computation of the same op on the same seeded inputs; bf16-operand kernels use inputs
already rounded to bf16 so the only differences are accumulation order and the final rounding."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from audioeditingcode_b200._lib import operand_torch_dtype
BF = operand_torch_dtype()        # the library build's 16-bit operand type (fp16 default, bf16 with AEDIT_OPERANDS=bf16)


@pytest.fixture(scope="module")
def ops():
    from audioeditingcode_b200.ops import CudaOps
    o = CudaOps()
    assert o.lib.ae_device_ok() == 1, "these tests need an sm_100 device"
    return o


def rnd(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def relerr(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 0), (256, 256, 512, 128), (300, 200, 200, 64), (64, 40, 72, 32),
                                      (4096, 128, 1152, 0), (128, 640, 5760, 0), (2, 8320, 1024, 0)])
def test_gemm_plain(ops, M, N, K, bn):
    A = rnd((M, K), 1, dtype=BF)
    W = rnd((N, K), 2, 1 / math.sqrt(K), dtype=BF)
    bias = rnd((N,), 3)
    res = rnd((M, N), 4)
    rowbias = rnd(((M + 15) // 16, N), 5)
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=BF)
    ops.gemm(A, W, out_f32=o32, out_bf16=o16, bias=bias, rowbias=rowbias, rows_per_group=16, residual=res, force_bn=bn)
    ref = A.float() @ W.float().t() + bias + res + rowbias.repeat_interleave(16, 0)[:M]
    torch.cuda.synchronize()
    assert relerr(o32, ref) < 2e-5
    assert relerr(o16, ref) < 4e-3
    # SiLU epilogue, no extras
    ops.gemm(A, W, out_f32=o32, act=1, force_bn=bn)
    assert relerr(o32, F.silu(A.float() @ W.float().t())) < 2e-5


@pytest.mark.parametrize("M,N,K,bn,rpg", [(128, 128, 64, 0, 16), (256, 256, 512, 128, 64), (300, 200, 200, 64, 16),
                                          (64, 40, 72, 32, 8), (4096, 128, 1152, 0, 4096), (2048, 384, 384, 0, 1024),
                                          (512, 576, 576, 0, 256), (130, 960, 960, 0, 64), (1000, 192, 192, 0, 100)])
@pytest.mark.parametrize("stages", [3, 6])
def test_gemm_epilogue_paths_bitexact(ops, M, N, K, bn, rpg, stages):
    """The coalesced shared-memory-staged epilogue (default) and the row-per-thread epilogue apply the same terms in
    the same order: identical bits, for both pipeline depths (the deep variant preloads the whole residual strip)."""
    A = rnd((M, K), 1, dtype=BF)
    W = rnd((N, K), 2, 1 / math.sqrt(K), dtype=BF)
    bias, res = rnd((N,), 3), rnd((M, N), 4)
    rowbias = rnd(((M + rpg - 1) // rpg, N), 5)
    outs = []
    for fast in (2, 0):
        ops.lib.ae_set_fast_epilogue(fast)
        try:
            o32 = torch.zeros(M, N, device="cuda")
            o16 = torch.zeros(M, N, device="cuda", dtype=BF)
            ops.gemm(A, W, out_f32=o32, out_bf16=o16, bias=bias, rowbias=rowbias, rows_per_group=rpg, residual=res,
                     act=1, alpha=0.7, force_bn=bn, force_split=1, force_stages=stages)
            outs.append((o32, o16))
        finally:
            ops.lib.ae_set_fast_epilogue(1)
    ref = F.silu(0.7 * (A.float() @ W.float().t()) + bias + res + rowbias.repeat_interleave(rpg, 0)[:M])
    assert relerr(outs[0][0], ref) < 2e-5
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("M,N,K,S", [(128, 960, 8640, 0), (128, 960, 8640, 9), (256, 576, 5184, 4), (100, 200, 1000, 3),
                                     (128, 7680, 960, 0), (512, 576, 2304, 0), (128, 960, 960, 5), (130, 960, 1920, 7),
                                     (128, 64, 4096, 32)])
def test_gemm_splitk(ops, M, N, K, S):
    """split-K (fp32 partials + fixed-order reduce kernel) gives the same result as the single-pass kernel up to
    fp32 summation order, with every epilogue term applied exactly once."""
    A = rnd((M, K), 1, dtype=BF)
    W = rnd((N, K), 2, 1 / math.sqrt(K), dtype=BF)
    bias, res = rnd((N,), 3), rnd((M, N), 4)
    rowbias = rnd(((M + 63) // 64, N), 5)
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=BF)
    ops.gemm(A, W, out_f32=o32, out_bf16=o16, bias=bias, rowbias=rowbias, rows_per_group=64, residual=res, act=1,
             alpha=0.5, force_split=S)
    ref = F.silu(0.5 * (A.float() @ W.float().t()) + bias + res + rowbias.repeat_interleave(64, 0)[:M])
    assert relerr(o32, ref) < 2e-5
    assert relerr(o16, ref) < 4e-3
    o_ns = torch.empty_like(o32)
    ops.gemm(A, W, out_f32=o_ns, bias=bias, rowbias=rowbias, rows_per_group=64, residual=res, act=1, alpha=0.5,
             force_split=1)
    assert relerr(o32, o_ns) < 1e-5
    # determinism: same call, same bits
    o_b = torch.zeros_like(o32)
    o_h = torch.zeros_like(o16)
    ops.gemm(A, W, out_f32=o_b, out_bf16=o_h, bias=bias, rowbias=rowbias, rows_per_group=64, residual=res, act=1,
             alpha=0.5, force_split=S)
    assert torch.equal(o32, o_b) and torch.equal(o16, o_h)


@pytest.mark.parametrize("M,N,K,rows,S,bn", [(128, 960, 960, 64, 1, 0), (128, 960, 8640, 64, 0, 0), (512, 576, 2304, 256, 4, 0),
                                             (2048, 384, 384, 1024, 1, 32), (8192, 192, 1728, 4096, 1, 64),
                                             (8192, 192, 1728, 4096, 1, 128), (256, 200, 512, 32, 3, 0),
                                             (96, 960, 960, 32, 1, 0), (6400, 384, 384, 64, 1, 0)])
def test_gemm_column_statistics(ops, M, N, K, rows, S, bn):
    """ae_gemm_args.colstats: fixed-point per-(sample, column) sum / sum of squares of the final fp32 output, from the
    staged epilogue (S == 1) or the split-K reduce kernel; the output bits are those of the plain call; accumulating
    twice doubles the integers exactly (order-independent integer atomics => deterministic)."""
    A = rnd((M, K), 1, dtype=BF)
    W = rnd((N, K), 2, 1 / math.sqrt(K), dtype=BF)
    bias, res = rnd((N,), 3), rnd((M, N), 4)
    rowbias = rnd((M // rows, N), 5)
    o_ref = torch.empty(M, N, device="cuda")
    ops.gemm(A, W, out_f32=o_ref, bias=bias, rowbias=rowbias, rows_per_group=rows, residual=res, force_split=S,
             force_bn=bn)
    nb = M // rows
    cs = torch.zeros(nb, N, 2, dtype=torch.int64, device="cuda")
    o = torch.empty_like(o_ref)
    ops.gemm(A, W, out_f32=o, bias=bias, rowbias=rowbias, rows_per_group=rows, residual=res, force_split=S, force_bn=bn,
             colstats=cs, cs_rows=rows)
    assert torch.equal(o, o_ref)
    x = o.double().view(nb, rows, N)
    su = cs[..., 0].double() / 2 ** 28
    sq = cs[..., 1].double() / 2 ** 24
    assert (su - x.sum(1)).abs().max().item() < 1e-3 * max(1.0, x.sum(1).abs().max().item())
    assert ((sq - (x * x).sum(1)).abs() / (x * x).sum(1)).max().item() < 1e-5
    cs2 = cs.clone()
    ops.gemm(A, W, out_f32=o, bias=bias, rowbias=rowbias, rows_per_group=rows, residual=res, force_split=S, force_bn=bn,
             colstats=cs2, cs_rows=rows)
    assert torch.equal(cs2, 2 * cs)


@pytest.mark.parametrize("B,HW,C1,C2,silu,stream", [(2, 4096, 192, 0, True, 0), (2, 1024, 576, 384, True, 0),
                                                    (2, 64, 960, 960, False, 0), (3, 256, 960, 576, True, 1),
                                                    (5, 1024, 384, 0, True, 1)])
def test_groupnorm_from_column_statistics(ops, B, HW, C1, C2, silu, stream):
    """ae_groupnorm_cs (statistics from the producers' column sums, one launch) agrees with ae_groupnorm (statistics
    pass over the tensor) to fp32 rounding of mean / rstd, in both apply kernels."""
    x1 = rnd((B, HW, C1), 1) + 0.5
    x2 = rnd((B, HW, C2), 2, 2.0) if C2 else None
    C = C1 + C2
    gamma, beta = rnd((C,), 3) * 0.1 + 1, rnd((C,), 4) * 0.1

    def colsums(x):
        xd = x.double()
        return torch.stack([(xd.sum(1) * 2 ** 28).round(), ((xd * xd).sum(1) * 2 ** 24).round()], -1).to(torch.int64).contiguous()
    cs1 = colsums(x1)
    cs2 = colsums(x2) if C2 else None
    ops.lib.ae_set_gn_stream_min_bytes(0 if stream else 1 << 60)
    try:
        ref = torch.zeros(B, HW, C, device="cuda", dtype=BF)
        ops.groupnorm(x1, x2, gamma, beta, 1e-5, 32, silu, ref)
        out = torch.zeros_like(ref)
        raw = torch.zeros_like(ref)
        ops.groupnorm(x1, x2, gamma, beta, 1e-5, 32, silu, out, raw_out=raw, cs1=cs1, cs2=cs2)
    finally:
        ops.lib.ae_set_gn_stream_min_bytes(8 << 20)
    d = (out.float() - ref.float()).abs()
    # identical up to a bf16 ulp where the fp32 mean / rstd differ in their last bit
    assert (d > 0).float().mean().item() < 0.02 and d.max().item() <= 0.04
    x = x1 if x2 is None else torch.cat([x1, x2], -1)
    assert torch.equal(raw, x.to(BF))


@pytest.mark.parametrize("M,N,K,kind", [
    (128 * 160 + 37, 384, 384, "res"),        # > one tile per SM, ragged M tail, residual + bias + fp32/bf16 out (staged)
    (128 * 300, 192, 1728, "rowbias"),        # BN = 64 tiles, time-embedding row bias (staged, multi-wave fp32 out)
    (128 * 170, 1152, 384, "bf16"),           # bf16-only output (row-per-thread epilogue)
    (128 * 40, 200, 72, "res"),               # K shorter than the ring, ragged N
    (128 * 64, 1536, 384, "geglu"),           # GEGLU epilogue
    (128 * 64, 384, 384, "stats"),            # GroupNorm column statistics from the staged epilogue
    (128 * 3, 960, 960, "res"),               # fewer tiles than SMs
])
def test_gemm_persistent_same_bits(ops, M, N, K, kind):
    """The persistent kernel (one CTA per SM walking the tiles, TMEM accumulator double-buffered) produces the bits of the
    one-CTA-per-tile kernel for every epilogue."""
    A = rnd((M, K), 1, dtype=BF)
    W = rnd((N, K), 2, 1 / math.sqrt(K), dtype=BF)
    bias = rnd((N,), 3)
    outs = []
    for fp in (-1, 1, 1):
        kw = dict(force_persistent=fp, force_split=1)
        if kind == "res":
            res = rnd((M, N), 4)
            o32 = torch.zeros(M, N, device="cuda")
            o16 = torch.zeros(M, N, device="cuda", dtype=BF)
            ops.gemm(A, W, out_f32=o32, out_bf16=o16, bias=bias, residual=res, **kw)
            outs.append((o32, o16))
        elif kind == "rowbias":
            rb = rnd((M // 128 // 4 + 1, N), 5)
            o32 = torch.zeros(M, N, device="cuda")
            ops.gemm(A, W, out_f32=o32, bias=bias, rowbias=rb, rows_per_group=512, **kw)
            outs.append((o32,))
        elif kind == "bf16":
            o16 = torch.zeros(M, N, device="cuda", dtype=BF)
            ops.gemm(A, W, out_bf16=o16, **kw)
            outs.append((o16,))
        elif kind == "geglu":
            o16 = torch.zeros(M, N // 2, device="cuda", dtype=BF)
            ops.gemm(A, W, out_bf16=o16, bias=bias, act=2, **kw)
            outs.append((o16,))
        elif kind == "stats":
            res = rnd((M, N), 4)
            o32 = torch.zeros(M, N, device="cuda")
            cs = torch.zeros(M // 1024, N, 2, dtype=torch.int64, device="cuda")
            ops.gemm(A, W, out_f32=o32, bias=bias, residual=res, colstats=cs, cs_rows=1024, **kw)
            outs.append((o32, cs))
    for a, b, c in zip(*outs):
        assert torch.equal(a, b) and torch.equal(b, c)
    if kind == "res":
        ref = A.float() @ W.float().t() + bias + res
        assert relerr(outs[1][0], ref) < 2e-5





def test_gemm_persistent_implicit_conv(ops):
    B, H, Wd, C, Co = 24, 64, 16, 192, 192
    x = rnd((B, H, Wd, C), 1, dtype=BF)
    wt = rnd((Co, C, 3, 3), 2, 1 / math.sqrt(C * 9), dtype=BF)
    bias = rnd((Co,), 3)
    Wp = wt.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    res = rnd((B * H * Wd, Co), 4)
    outs = []
    for fp in (-1, 1):
        o = torch.zeros(B * H * Wd, Co, device="cuda")
        ops.gemm(x, Wp, out_f32=o, bias=bias, residual=res, conv=(B, H, Wd, C, 3, 3, 1, 1), force_persistent=fp, force_split=1)
        outs.append(o)
    assert torch.equal(outs[0], outs[1])
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Co) + res
    assert relerr(outs[1], ref) < 2e-5


@pytest.mark.parametrize("M,C", [(300, 64), (4096, 192), (128, 960)])
def test_gemm_geglu_epilogue(ops, M, C):
    """FF1 + GEGLU fused (act=2, interleaved weight rows) == chunk(2) -> value * gelu(gate) (attention.py:37-44)."""
    inner = 4 * C
    x = rnd((M, C), 1, dtype=BF)
    Wf = rnd((2 * inner, C), 2, 1 / math.sqrt(C), dtype=BF)
    bf = rnd((2 * inner,), 3)
    idx = torch.arange(inner).view(-1, 16)
    perm = torch.cat([idx, idx + inner], 1).reshape(-1).cuda()
    out = torch.empty(M, inner, device="cuda", dtype=BF)
    ops.gemm(x, Wf[perm].contiguous(), out_bf16=out, bias=bf[perm].contiguous(), act=2)
    h = x.float() @ Wf.float().t() + bf
    a, g = h.chunk(2, -1)
    assert relerr(out, a * F.gelu(g)) < 4e-3
    # row-per-thread epilogue: same bits
    out2 = torch.empty_like(out)
    ops.lib.ae_set_fast_epilogue(2)
    try:
        ops.gemm(x, Wf[perm].contiguous(), out_bf16=out2, bias=bf[perm].contiguous(), act=2)
    finally:
        ops.lib.ae_set_fast_epilogue(1)
    assert torch.equal(out, out2)


def test_gemm_batched(ops):
    Bz, M, N, K = 3, 200, 96, 160
    A = rnd((Bz, M, K), 1, dtype=BF)
    W = rnd((Bz, N, K), 2, 0.1, dtype=BF)
    o = torch.empty(Bz, M, N, device="cuda")
    ops.gemm(A, W, out_f32=o, batch=Bz, strideA=M * K, strideW=N * K, stride_out=M * N, alpha=0.5)
    ref = 0.5 * torch.einsum("bmk,bnk->bmn", A.float(), W.float())
    assert relerr(o, ref) < 2e-5


@pytest.mark.parametrize("B,H,W,C,Co,kh,kw,dh,dw", [(2, 16, 16, 64, 128, 3, 3, 1, 1), (1, 256, 16, 128, 128, 3, 3, 1, 1),
                                                    (2, 32, 2, 640, 640, 3, 3, 1, 1), (3, 64, 4, 384, 384, 3, 3, 1, 1),
                                                    (1, 1, 1024, 64, 64, 1, 7, 1, 3), (2, 8, 8, 192, 8, 3, 3, 1, 1),
                                                    (2, 128, 8, 256, 256, 1, 1, 1, 1)])
def test_gemm_implicit_conv(ops, B, H, W, C, Co, kh, kw, dh, dw):
    assert ops.conv_supported(B, H, W, C)
    x = rnd((B, H, W, C), 1, dtype=BF)
    wt = rnd((Co, C, kh, kw), 2, 1 / math.sqrt(C * kh * kw), dtype=BF)
    bias = rnd((Co,), 3)
    temb = rnd((B, Co), 4)
    res = rnd((B * H * W, Co), 5)
    Wp = wt.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    out = torch.empty(B * H * W, Co, device="cuda")
    ops.gemm(x, Wp, out_f32=out, bias=bias, rowbias=temb, rows_per_group=H * W, residual=res,
             conv=(B, H, W, C, kh, kw, dh, dw))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=(dh * (kh - 1) // 2, dw * (kw - 1) // 2),
                   dilation=(dh, dw))
    ref = ref + temb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Co) + res
    assert relerr(out, ref) < 2e-5


@pytest.mark.parametrize("stride,pad,H,W,C", [(2, 1, 16, 16, 64), (1, 1, 20, 16, 8), (2, 0, 17, 9, 32)])
def test_im2col(ops, stride, pad, H, W, C):
    B, kh, kw = 2, 3, 3
    x = rnd((B, H, W, C), 1)
    Ho = (H + 2 * pad - 3) // stride + 1
    Wo = (W + 2 * pad - 3) // stride + 1
    K = 9 * C
    ld = (K + 7) // 8 * 8
    col = torch.empty(B * Ho * Wo, ld, device="cuda", dtype=BF)
    ops.im2col(x, B, H, W, C, kh, kw, stride, 1, pad, pad, Ho, Wo, col)
    cols = F.unfold(x.permute(0, 3, 1, 2), 3, padding=pad, stride=stride)
    cols = cols.reshape(B, C, 9, -1).permute(0, 3, 2, 1).reshape(B * Ho * Wo, K)
    assert torch.equal(col[:, :K], cols.to(BF))


@pytest.mark.parametrize("B,HW,C1,C2,G,silu", [(2, 4096, 128, 0, 32, True), (3, 64, 640, 640, 32, True),
                                               (2, 256, 192, 96, 32, False), (1, 1000, 576, 384, 32, True),
                                               (2, 4096, 192, 0, 32, True), (2, 1024, 576, 384, 32, True),
                                               (12, 256, 192, 0, 32, True), (1, 7, 64, 0, 32, True)])
def test_groupnorm(ops, B, HW, C1, C2, G, silu):
    """statistics launch + apply launch (the path without producer column statistics)."""
    _groupnorm_case(ops, B, HW, C1, C2, G, silu)


def test_groupnorm_batch_independent(ops):
    """A sample's GroupNorm bits do not depend on the batch it is in at the batch sizes of the reverse process (1 or 2
    rows: the position slices per sample are the same, and they are reduced in a fixed shape)."""
    x = rnd((2, 1024, 384), 7) * 3 + 0.2
    gamma, beta = rnd((384,), 3) * 0.1 + 1, rnd((384,), 4) * 0.1
    o2 = torch.empty(2, 1024, 384, device="cuda", dtype=BF)
    o1 = torch.empty(1, 1024, 384, device="cuda", dtype=BF)
    ops.groupnorm(x, None, gamma, beta, 1e-5, 32, True, o2)
    for b in range(2):
        ops.groupnorm(x[b:b + 1].contiguous(), None, gamma, beta, 1e-5, 32, True, o1)
        assert torch.equal(o1[0], o2[b])
    # same sample, different neighbours
    y = x.clone()
    y[1] = rnd((1024, 384), 9)
    o2b = torch.empty_like(o2)
    ops.groupnorm(y, None, gamma, beta, 1e-5, 32, True, o2b)
    assert torch.equal(o2b[0], o2[0])


@pytest.mark.parametrize("B,HW,C1,C2,silu", [(3, 1000, 192, 0, True), (2, 4096, 384, 192, True), (5, 64, 960, 960, False),
                                             (2, 256, 960, 576, True)])
def test_groupnorm_stream_apply_same_bits(ops, B, HW, C1, C2, silu):
    """The streaming apply kernel (large tensors: forward-process chunks) and the one-round apply kernel (small
    tensors: reverse process) produce identical bits, including the raw bf16 copy and the concatenated input."""
    x1 = rnd((B, HW, C1), 1) + 0.5
    x2 = rnd((B, HW, C2), 2, 2.0) if C2 else None
    C = C1 + C2
    gamma, beta = rnd((C,), 3) * 0.1 + 1, rnd((C,), 4) * 0.1
    outs = []
    try:
        for min_bytes in (1 << 60, 0):
            ops.lib.ae_set_gn_stream_min_bytes(min_bytes)
            out = torch.zeros(B, HW, C, device="cuda", dtype=BF)
            raw = torch.zeros(B, HW, C, device="cuda", dtype=BF)
            ops.groupnorm(x1, x2, gamma, beta, 1e-5, 32, silu, out, raw_out=raw)
            outs.append((out, raw))
    finally:
        ops.lib.ae_set_gn_stream_min_bytes(8 << 20)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    x = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert (outs[1][0].float() - ref).abs().max().item() < 0.05


def _groupnorm_case(ops, B, HW, C1, C2, G, silu):
    x1 = rnd((B, HW, C1), 1) + 0.5
    x2 = rnd((B, HW, C2), 2, 2.0) if C2 else None
    C = C1 + C2
    gamma, beta = rnd((C,), 3) * 0.1 + 1, rnd((C,), 4) * 0.1
    out = torch.empty(B, HW, C, device="cuda", dtype=BF)
    raw = torch.empty(B, HW, C, device="cuda", dtype=BF)
    ops.groupnorm(x1, x2, gamma, beta, 1e-5, G, silu, out, raw_out=raw)
    x = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.group_norm(x.permute(0, 2, 1), G, gamma, beta, 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert (out.float() - ref).abs().max().item() < 3e-2
    assert relerr(out, ref) < 4e-3
    assert torch.equal(raw, x.to(BF))
    # second call reuses the self-re-arming workspace
    ops.groupnorm(x1, x2, gamma, beta, 1e-5, G, silu, out)
    assert relerr(out, ref) < 4e-3


def test_layernorm_geglu(ops):
    x = rnd((300, 384), 1) * 2 + 0.3
    g, b = rnd((384,), 2) * 0.1 + 1, rnd((384,), 3) * 0.1
    out = torch.empty(300, 384, device="cuda", dtype=BF)
    ops.layernorm(x, g, b, out)
    assert relerr(out, F.layer_norm(x, (384,), g, b)) < 4e-3
    h = rnd((100, 2 * 256), 4, dtype=BF)
    o = torch.empty(100, 256, device="cuda", dtype=BF)
    ops.geglu(h, o)
    a, gt = h.float().chunk(2, -1)
    assert relerr(o, a * F.gelu(gt)) < 4e-3


@pytest.mark.parametrize("d,heads,Tq,Tk", [(32, 4, 64, 64), (32, 8, 1024, 1024), (64, 5, 200, 200), (48, 8, 256, 256),
                                           (72, 8, 64, 64), (120, 8, 64, 64), (160, 8, 64, 64), (96, 8, 130, 77)])
def test_attention_self(ops, d, heads, Tq, Tk):
    B, C = 2, heads * d
    q = rnd((B, Tq, C), 1, dtype=BF)
    k = rnd((B, Tk, C), 2, dtype=BF)
    v = rnd((B, Tk, C), 3, dtype=BF)
    out = torch.empty(B * Tq, C, device="cuda", dtype=BF)
    ops.attention(q, k, v, out, heads, d, d ** -0.5, Tq, Tk, B, C, Tq * C, C, Tk * C, C, Tk * C)
    qh, kh, vh = (t.float().view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    ref = (qh @ kh.transpose(-1, -2) * d ** -0.5).softmax(-1) @ vh
    ref = ref.transpose(1, 2).reshape(B * Tq, C)
    assert relerr(out, ref) < 1e-2


@pytest.mark.parametrize("B,d,heads,Tq,Tk,bias", [
    (2, 48, 8, 1024, 1024, False),     # AudioLDM2-large level 1, reverse step (NT = 1, DP = 64)
    (2, 64, 5, 1024, 1024, False),     # TANGO head dim
    (1, 32, 4, 256, 256, False),
    (2, 72, 8, 256, 256, False),       # DP = 128 (d padded 72 -> 80 by TMA zero fill), BKEYS = 64
    (2, 120, 8, 128, 192, False),      # d = 120 -> 128, partial last key block
    (3, 48, 8, 1000, 1000, False),     # ragged query tile and ragged key block
    (2, 48, 4, 384, 333, True),        # additive key bias (+ ragged keys)
    (40, 48, 8, 128, 128, False),      # 320 tiles -> NT = 2 (two query tiles per CTA)
    (20, 80, 8, 256, 256, True),       # NT = 2, DP = 128, bias
])
def test_attention_tcgen05(ops, B, d, heads, Tq, Tk, bias):
    """tcgen05 / TMEM / TMA attention (csrc/attn_tc.cu) vs a plain PyTorch fp32 softmax attention on the same operand-
    rounded inputs, and vs the mma.sync kernel (ae_set_attention_tc(0)); deterministic (same call twice: same bits)."""
    C = heads * d
    q = rnd((B, Tq, 3 * C), 1, dtype=BF)                       # q | k | v packed like the fused qkv projection's output
    kk = q if Tk == Tq else rnd((B, Tk, 3 * C), 2, dtype=BF)
    kb = None
    if bias:
        kb = torch.zeros(B, Tk, device="cuda")
        kb[:, Tk - 7:] = -10000.0
        kb[0, 3] = -2.5
    args = (heads, d, d ** -0.5, Tq, Tk, B, 3 * C, Tq * 3 * C, 3 * C, Tk * 3 * C, 3 * C, Tk * 3 * C)
    outs = []
    try:
        for tc in (1, 1, 0, 2, 3, 4, 5):  # auto twice, mma.sync, then the forced variants (row split / two tiles / plain / both)
            ops.lib.ae_set_attention_tc(tc)
            o = torch.zeros(B * Tq, C, device="cuda", dtype=BF)
            ops.attention(q, kk[..., C:], kk[..., 2 * C:], o, *args, bias=kb)
            outs.append(o)
    finally:
        ops.lib.ae_set_attention_tc(1)
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[4], outs[5])                      # one or two query tiles per CTA: same bits per row
    qh = q[..., :C].float().view(B, Tq, heads, d).transpose(1, 2)
    kh = kk[..., C:2 * C].float().view(B, Tk, heads, d).transpose(1, 2)
    vh = kk[..., 2 * C:].float().view(B, Tk, heads, d).transpose(1, 2)
    sc = qh @ kh.transpose(-1, -2) * d ** -0.5
    if kb is not None:
        sc = sc + kb[:, None, None, :]
    ref = (sc.softmax(-1) @ vh).transpose(1, 2).reshape(B * Tq, C)
    e_tc, e_old = relerr(outs[0], ref), relerr(outs[2], ref)
    print(f"B={B} d={d} T={Tq}/{Tk}: rel err tcgen05 {e_tc:.2e}  mma.sync {e_old:.2e}")
    assert e_tc < 1e-2 and relerr(outs[0], outs[2]) < 1e-2
    for o in outs[3:]:
        assert relerr(o, ref) < 1e-2


def test_attention_cross_masked(ops):
    B, heads, d, Tq, R, L = 4, 8, 48, 256, 2, 13
    C = heads * d
    q = rnd((B, Tq, C), 1, dtype=BF)
    kv = rnd((R, L, 2 * C), 2, dtype=BF)
    slot = torch.tensor([0, 1, 1, 0], dtype=torch.int32, device="cuda")
    mask = torch.ones(R, L, device="cuda")
    mask[1, 9:] = 0
    bias = (1 - mask) * -10000.0
    out = torch.empty(B * Tq, C, device="cuda", dtype=BF)
    ops.attention(q, kv, kv[:, :, C:], out, heads, d, d ** -0.5, Tq, L, B, C, Tq * C, 2 * C, L * 2 * C, 2 * C, L * 2 * C,
                  kv_map=slot, bias=bias)
    k = kv[:, :, :C][slot.long()].float().view(B, L, heads, d).transpose(1, 2)
    v = kv[:, :, C:][slot.long()].float().view(B, L, heads, d).transpose(1, 2)
    qh = q.float().view(B, Tq, heads, d).transpose(1, 2)
    s = qh @ k.transpose(-1, -2) * d ** -0.5 + bias[slot.long()][:, None, None, :]
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * Tq, C)
    assert relerr(out, ref) < 1e-2


def test_small_movers(ops):
    t = torch.tensor([981, 501, 1, 0], device="cuda")
    out = torch.empty(4, 128, device="cuda", dtype=BF)
    ops.timestep_embedding(t, 128, out)
    half = 64
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert (out.float() - ref).abs().max().item() < 1e-2
    x = rnd((2, 5, 7, 64), 1)
    up = torch.empty(2, 10, 13, 64, device="cuda", dtype=BF)
    ops.upsample_nearest(x, 2, 5, 7, 64, 10, 13, up)
    ref = F.interpolate(x.permute(0, 3, 1, 2), size=(10, 13), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up, ref.to(BF))
    y = rnd((3, 8, 20, 16), 2)
    nhwc = torch.empty(3, 20, 16, 8, device="cuda")
    ops.nchw_to_nhwc(y, out_f32=nhwc)
    assert torch.equal(nhwc, y.permute(0, 2, 3, 1).contiguous())
    back = torch.empty_like(y)
    ops.nhwc_to_nchw(nhwc, 3, 8, 20, 16, back)
    assert torch.equal(back, y)
    s = rnd((70, 333), 3)
    so = torch.empty(70, 333, device="cuda", dtype=BF)
    ops.softmax_rows(s, so)
    assert relerr(so, s.softmax(-1)) < 4e-3
    tb = rnd((2, 50, 70), 4, dtype=BF)
    to = torch.empty(2, 70, 50, device="cuda", dtype=BF)
    ops.transpose_bf16(tb, to)
    assert torch.equal(to, tb.transpose(1, 2).contiguous())
