"""Per-kernel cost microbenchmarks (GPU box): each case is R back-to-back launches captured in one CUDA graph,
timed with CUDA events over several replays.  Prints us/launch and, for GEMMs, TFLOP/s."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioeditingcode_b200.ops import CudaOps  # noqa: E402

ops = CudaOps()
from audioeditingcode_b200._lib import operand_torch_dtype
BF = operand_torch_dtype()        # the library build's 16-bit operand type (fp16 default, bf16 with AEDIT_OPERANDS=bf16)
dev = "cuda"


def bench(name, fn, R=100, reps=5, flops=0.0, launches_per_call=1):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1000 / (R * reps)
    extra = f"  {flops / us / 1e6:8.1f} TFLOP/s" if flops else ""
    print(f"{name:58s} {us / launches_per_call:8.2f} us/launch{extra}", flush=True)
    return us


def gemm_case(M, N, K, bn, conv=None, res=False, split=0):
    if conv is not None:
        B, H, W, C = conv
        A = torch.randn(B, H, W, C, device=dev).to(BF)
        K = 9 * C
    else:
        A = torch.randn(M, K, device=dev).to(BF)
    Wt = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF)
    out = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None

    def fn():
        ops.gemm(A, Wt, out_f32=out, bias=bias, residual=r, force_bn=bn, force_split=split,
                 conv=None if conv is None else (*conv, 3, 3, 1, 1))
    return fn, 2.0 * M * N * K


x = torch.randn(256, device=dev)
xb = torch.empty(256, device=dev, dtype=BF)
bench("tiny elementwise (cast 256 el)", lambda: ops.cast_bf16(x, xb))

ln_x = torch.randn(128, 960, device=dev)
ln_g, ln_b = torch.ones(960, device=dev), torch.zeros(960, device=dev)
ln_o = torch.empty(128, 960, device=dev, dtype=BF)
bench("layernorm rows=128 C=960", lambda: ops.layernorm(ln_x, ln_g, ln_b, ln_o))
ln_x2 = torch.randn(8192, 192, device=dev)
ln_o2 = torch.empty(8192, 192, device=dev, dtype=BF)
g192, b192 = torch.ones(192, device=dev), torch.zeros(192, device=dev)
bench("layernorm rows=8192 C=192", lambda: ops.layernorm(ln_x2, g192, b192, ln_o2))

for (B, HW, C) in [(2, 64, 960), (2, 4096, 192), (1, 4096, 192), (2, 4096, 384), (2, 1024, 576), (2, 1024, 960),
                   (2, 256, 1536), (16, 4096, 192), (16, 64, 960)]:
    gx = torch.randn(B, HW, C, device=dev)
    gg, gb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    go = torch.empty(B, HW, C, device=dev, dtype=BF)
    bench(f"groupnorm B={B} HW={HW} C={C} (stats + apply launches)",
          lambda: ops.groupnorm(gx, None, gg, gb, 1e-5, 32, True, go), launches_per_call=2)

for bn in (32, 64, 128):
    fn, fl = gemm_case(128, 960, 960, bn, split=1)
    bench(f"gemm M=128 N=960 K=960 bn={bn} (no split)", fn, flops=fl)
for bn in (32, 64, 128):
    fn, fl = gemm_case(128, 960, 8640, bn, split=1)
    bench(f"gemm M=128 N=960 K=8640 bn={bn} (no split)", fn, flops=fl)
for sp in (4, 9, 18):
    fn, fl = gemm_case(128, 960, 8640, 128, split=sp)
    bench(f"gemm M=128 N=960 K=8640 bn=128 ws-split={sp} (2 launches)", fn, flops=fl)
fn, fl = gemm_case(128, 960, 0, 128, conv=(2, 32, 2, 960), split=0)
bench("conv3x3 B=2 32x2 C=960->960 bn=128 split=auto (implicit)", fn, flops=fl)
fn, fl = gemm_case(512, 576, 0, 0, conv=(2, 64, 4, 576), split=0)
bench("conv3x3 B=2 64x4 C=576->576 auto (implicit)", fn, flops=fl)
fn, fl = gemm_case(512, 576, 5184, 0, split=0)
bench("gemm   M=512 N=576 K=5184 auto (explicit-gather equivalent)", fn, flops=fl)
fn, fl = gemm_case(128, 960, 0, 32, conv=(2, 32, 2, 960))
bench("conv3x3 B=2 32x2 C=960->960 bn=32 (implicit)", fn, flops=fl)
for bn in (64, 128):
    fn, fl = gemm_case(8192, 192, 0, bn, conv=(2, 256, 16, 192))
    bench(f"conv3x3 B=2 256x16 C=192->192 bn={bn}", fn, flops=fl)
for bn in (64, 128):
    fn, fl = gemm_case(65536, 192, 0, bn, conv=(16, 256, 16, 192), res=True)
    bench(f"conv3x3 B=16 256x16 C=192->192 bn={bn} (+res)", fn, R=20, flops=fl)
fn, fl = gemm_case(16384, 384, 0, 128, conv=(16, 128, 8, 384))
bench("conv3x3 B=16 128x8 C=384->384 bn=128", fn, R=20, flops=fl)
fn, fl = gemm_case(4096, 576, 0, 128, conv=(16, 64, 4, 576))
bench("conv3x3 B=16 64x4 C=576->576 bn=128", fn, R=20, flops=fl)
fn, fl = gemm_case(1024, 960, 0, 128, conv=(16, 32, 2, 960))
bench("conv3x3 B=16 32x2 C=960->960 bn=128", fn, R=20, flops=fl)
fn, fl = gemm_case(16384, 3072, 384, 128)
bench("linear M=16384 N=3072 K=384 (ff1 lvl1 B=16)", fn, R=20, flops=fl)
fn, fl = gemm_case(8192, 8192, 8192, 128)
bench("gemm 8192^3 bn=128", fn, R=3, reps=3, flops=fl)

# alternating small gemm + layernorm (shared-memory carve-out switching)
fn_g, fl = gemm_case(128, 960, 960, 32)


def alt():
    fn_g()
    ops.layernorm(ln_x, ln_g, ln_b, ln_o)
bench("alternating [gemm M=128 N=960 K=960 bn=32, layernorm]", alt, launches_per_call=2)

for (d, h, T, B) in [(120, 8, 64, 2), (72, 8, 256, 2), (48, 8, 1024, 2), (48, 8, 1024, 16)]:
    C = d * h
    q = torch.randn(B, T, 3 * C, device=dev).to(BF)
    o = torch.empty(B * T, C, device=dev, dtype=BF)
    bench(f"attention d={d} heads={h} T={T} B={B}",
          lambda: ops.attention(q, q[:, :, C:], q[:, :, 2 * C:], o, h, d, d ** -0.5, T, T, B, 3 * C, T * 3 * C, 3 * C,
                                T * 3 * C, 3 * C, T * 3 * C), flops=4.0 * B * h * T * T * d)
