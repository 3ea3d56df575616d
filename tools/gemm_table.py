"""Per-shape table of the ae_gemm calls of one U-Net evaluation: every distinct (M, N, K, conv, epilogue) signature is
replayed alone in a CUDA graph (20 back-to-back launches, warm caches) and reported with its call count, time per
call, TFLOP/s and algorithmic GB/s, sorted by its share of the evaluation's GEMM time.

    python tools/gemm_table.py [--config audioldm2-large-10s] [--batch 2 100]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="audioldm2-large-10s")
ap.add_argument("--batch", type=int, nargs="+", default=[2, 100])
ap.add_argument("--top", type=int, default=40)
args = ap.parse_args()
spec = bench.CONFIGS[args.config]
dev = torch.device("cuda", 0)
m, cfg = bench.build_model(spec, dev)
ops = m.engine.ops
from audioeditingcode_b200.ddm_inversion.inversion_utils import _loop_text  # noqa: E402
text, cl = _loop_text(m, [""], ["a recording of a dog barking"])


def sig(a, k):
    A, W = a[0], a[1]
    conv = k.get("conv")
    N = W.shape[-2]
    if conv is not None:
        B_, H_, W_, C_, kh, kw, dh, dw = conv
        M, K = B_ * H_ * W_, kh * kw * C_
    else:
        M = int(k["M"]) if k.get("M") is not None else A.shape[-2]
        K = int(k["K"]) if k.get("K") is not None else A.shape[-1]
    ep = "".join(c for c, on in (("b", k.get("bias") is not None), ("t", k.get("rowbias") is not None),
                                 ("r", k.get("residual") is not None), ("F", k.get("out_f32") is not None),
                                 ("H", k.get("out_bf16") is not None), ("g", k.get("act", 0) == 2),
                                 ("s", k.get("act", 0) == 1)) if on)
    return (M, N, K, "conv%dx%d" % (conv[4], conv[5]) if conv is not None else "lin", k.get("batch", 1), ep)


for B in args.batch:
    x = torch.randn(B, cfg.in_channels, spec["H"], spec["W"], device=dev)
    t = torch.full((B,), 501, dtype=torch.int64, device=dev)
    slot = (torch.arange(B, dtype=torch.int32, device=dev) % 2) if text is not None else None
    clb = None if cl is None else cl[(torch.arange(B, device=dev) % 2)]
    calls, keep = [], []
    orig, orig_empty = ops.gemm, ops.empty

    def rec(*a, **k):
        calls.append((a, k))
        orig(*a, **k)

    def keep_empty(*a, **k):
        tns = orig_empty(*a, **k)
        keep.append(tns)
        return tns
    m.engine.forward(x, t, text=text, slot_map=slot, class_labels=clb)
    ops.gemm, ops.empty = rec, keep_empty
    m.engine.forward(x, t, text=text, slot_map=slot, class_labels=clb)
    ops.gemm, ops.empty = orig, orig_empty
    torch.cuda.synchronize()
    groups = {}
    for a, k in calls:
        groups.setdefault(sig(a, k), []).append((a, k))
    rows = []
    for s, lst in groups.items():
        a, k = lst[0]
        reps = 20
        g = torch.cuda.CUDAGraph()
        orig(*a, **k)
        torch.cuda.synchronize()
        l0 = ops.launch_count()
        with torch.cuda.graph(g):
            for _ in range(reps):
                orig(*a, **k)
        nl = (ops.launch_count() - l0) // reps
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (3 * reps)
        M, N, K, kind, batch, ep = s
        flop = 2.0 * M * N * K * batch
        # algorithmic bytes: A once (conv: the image once), W once, outputs, residual
        a_bytes = (M * (K if kind == "lin" else K // (int(kind[4]) * int(kind[6]))) * 2) * batch
        byts = a_bytes + N * K * 2 * batch + M * N * batch * ((4 if "F" in ep else 0) + (2 if "H" in ep else 0) +
                                                             (4 if "r" in ep else 0))
        if "g" in ep:
            byts -= M * N * batch  # GEGLU epilogue writes N/2 bf16 columns
        rows.append((us * len(lst), len(lst), s, us, flop / us / 1e6, byts / us / 1e3, nl))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"== B={B}: {len(calls)} ae_gemm calls, {len(rows)} distinct shapes, sum of isolated times {tot/1e3:.2f} ms")
    print(f"{'M':>8} {'N':>5} {'K':>6} {'kind':>8} {'bt':>3} {'epi':>6} {'n':>4} {'us/call':>9} {'TFLOP/s':>8} {'GB/s':>7} "
          f"{'lnch':>4} {'share%':>6}")
    for tt, n, s, us, tf, gb, nl in rows[:args.top]:
        M, N, K, kind, batch, ep = s
        print(f"{M:8d} {N:5d} {K:6d} {kind:>8} {batch:3d} {ep:>6} {n:4d} {us:9.2f} {tf:8.1f} {gb:7.0f} {nl:4d} "
              f"{100*tt/tot:6.1f}")
    del keep, calls, groups
    torch.cuda.empty_cache()
