"""Profiling target for ncu (run under `ncu --profile-from-start off ...`): builds the benchmark model, warms up,
then executes — inside a cudaProfilerStart/Stop window, eagerly (no CUDA graph, so every kernel is a separate
launch ncu can attribute) — one CFG denoising step at B=2 (reverse-process shape) and one timestep-batched forward
chunk (B = 2*forward_batch)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="audioldm2-large-10s")
ap.add_argument("--forward-batch", type=int, default=8)
ap.add_argument("--only", default="both", choices=["both", "b2", "chunk"])
args = ap.parse_args()
spec = bench.CONFIGS[args.config]
dev = torch.device("cuda", 0)
m, cfg = bench.build_model(spec, dev)
from audioeditingcode_b200.ddm_inversion.inversion_utils import _loop_text  # noqa: E402
text, cl = _loop_text(m, [""], ["a recording of a dog barking"])


def step(B):
    x = torch.randn(B, cfg.in_channels, spec["H"], spec["W"], device=dev)
    t = torch.full((B,), 501, dtype=torch.int64, device=dev)
    slot = (torch.arange(B, dtype=torch.int32, device=dev) % 2) if text is not None else None
    c = None if cl is None else cl[(torch.arange(B, device=dev) % 2)]
    return m.engine.forward(x, t, text=text, slot_map=slot, class_labels=c)


for B in (2, 2 * args.forward_batch):
    step(B)
    step(B)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if args.only in ("both", "b2"):
    step(2)
if args.only in ("both", "chunk"):
    step(2 * args.forward_batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled window done; kernels launched so far:", m.engine.ops.launch_count())
