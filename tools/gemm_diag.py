"""Bring-up diagnostic for the tcgen05 GEMM (run on the GPU box): structured operands that make descriptor /
swizzle / pipeline mistakes visible as patterns instead of noise.  Prints one line per case."""
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioeditingcode_b200.ops import CudaOps  # noqa: E402

ops = CudaOps()
from audioeditingcode_b200._lib import operand_torch_dtype
BF = operand_torch_dtype()        # the library build's 16-bit operand type (fp16 default, bf16 with AEDIT_OPERANDS=bf16)


def run(M, N, K, bn, pattern):
    if pattern == "ident":
        A = torch.zeros(M, K)
        A[torch.arange(M), torch.arange(M) % K] = 1
        W = (torch.arange(N)[:, None] * 4 + (torch.arange(K)[None, :] % 4)).float() % 251
    elif pattern == "ones":
        A = torch.ones(M, K)
        W = torch.ones(N, K)
    else:
        g = torch.Generator().manual_seed(0)
        A = torch.randint(-3, 4, (M, K), generator=g).float()
        W = torch.randint(-3, 4, (N, K), generator=g).float()
    A, W = A.to(BF).cuda(), W.to(BF).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(A, W, out_f32=out, force_bn=bn)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    bad = ~torch.isclose(out, ref, atol=1e-3, rtol=1e-3)
    nbad = int(bad.sum())
    msg = f"M={M:5d} N={N:5d} K={K:5d} bn={bn:3d} {pattern:6s} bad={nbad}/{M*N}"
    if nbad:
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()[:8].tolist()
        msg += f" bad_rows[:8]={rows} bad_cols[:8]={cols} nan={int(torch.isnan(out).sum())}"
        r, c = rows[0], cols[0]
        msg += f"\n    out[{r},{c}:{c+6}]={out[r, c:c+6].tolist()} ref={ref[r, c:c+6].tolist()}"
    print(msg, flush=True)
    return nbad == 0


ok = True
for pattern in ("ones", "ident", "rand"):
    for (M, N, K, bn) in [(128, 32, 64, 32), (128, 64, 64, 64), (128, 128, 64, 128), (128, 128, 128, 128),
                          (128, 128, 1024, 128), (256, 256, 256, 128), (100, 50, 72, 64), (1000, 200, 320, 0)]:
        ok &= run(M, N, K, bn, pattern)
print("GEMM_DIAG", "PASS" if ok else "FAIL")
