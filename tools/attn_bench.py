"""Isolated timing of the attention kernel (CUDA graph of 20 back-to-back launches) per split-KV setting."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioeditingcode_b200.ops import CudaOps  # noqa: E402

ops = CudaOps()
from audioeditingcode_b200._lib import operand_torch_dtype
BF = operand_torch_dtype()        # the library build's 16-bit operand type (fp16 default, bf16 with AEDIT_OPERANDS=bf16)
for (B, heads, d, T) in [(2, 8, 48, 1024), (2, 8, 72, 256), (2, 8, 120, 64), (100, 8, 48, 1024)]:
    C = heads * d
    qkv = torch.randn(B * T, 3 * C, device="cuda").to(BF)
    out = torch.empty(B * T, C, device="cuda", dtype=BF)
    for ns in ([1, 2, 3, 4, 8] if B == 2 else [1]):
        ops.lib.ae_set_attention_split(ns)

        def run():
            ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], out, heads, d, d ** -0.5, T, T, B, 3 * C, T * 3 * C, 3 * C,
                          T * 3 * C, 3 * C, T * 3 * C)
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                run()
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / 100
        fl = 4.0 * B * heads * T * T * d
        print(f"attention B={B} heads={heads} d={d} T={T} split={ns}: {us:8.2f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
ops.lib.ae_set_attention_split(1)
