"""Isolated timing of the attention kernels (CUDA graph of 20 back-to-back launches): the tcgen05 / TMEM / TMA kernel
(csrc/attn_tc.cu) vs the mma.sync kernel (csrc/attn_kernels.cu, ae_set_attention_tc(0)) on the self-attention shapes of
the benchmarked U-Nets (AudioLDM2-large levels 1-2, TANGO level 0-1, the 30 s clip) at reverse-step and forward-chunk
batch sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioeditingcode_b200.ops import CudaOps  # noqa: E402
from audioeditingcode_b200._lib import operand_torch_dtype  # noqa: E402

ops = CudaOps()
BF = operand_torch_dtype()
SHAPES = [(2, 8, 48, 1024), (8, 8, 48, 1024), (100, 8, 48, 1024), (2, 8, 72, 256), (100, 8, 72, 256), (2, 5, 64, 4096),
          (20, 5, 64, 4096), (100, 10, 64, 1024), (2, 8, 32, 3072)]
for (B, heads, d, T) in SHAPES:
    C = heads * d
    qkv = torch.randn(B * T, 3 * C, device="cuda").to(BF)
    out = torch.empty(B * T, C, device="cuda", dtype=BF)
    res = {}
    for tc in (1, 0, 2, 3, 4, 5):
        ops.lib.ae_set_attention_tc(tc)

        def run():
            ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], out, heads, d, d ** -0.5, T, T, B, 3 * C, T * 3 * C, 3 * C,
                          T * 3 * C, 3 * C, T * 3 * C)
        run()
        torch.cuda.synchronize()
        n = 20 if B * T <= 32768 else 4
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                run()
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        res[tc] = s.elapsed_time(e) * 1e3 / (5 * n)
    fl = 4.0 * B * heads * T * T * d
    print(f"attention B={B:3d} heads={heads:2d} d={d:3d} T={T:4d}: tcgen05 {res[1]:9.2f} us {fl / res[1] / 1e6:7.1f} TFLOP/s | "
          f"mma.sync {res[0]:9.2f} us {fl / res[0] / 1e6:7.1f} TFLOP/s | speed-up {res[0] / res[1]:.2f}x | forced variants: "
          f"row-split {res[2]:.1f} us, two-tiles {res[3]:.1f} us, plain {res[4]:.1f} us, two-tiles+row-split {res[5]:.1f} us", flush=True)
ops.lib.ae_set_attention_tc(1)
