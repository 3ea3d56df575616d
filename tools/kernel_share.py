"""Marginal cost of each kernel family INSIDE the captured U-Net graph (leave-one-family-out): the graph is
re-captured with the family's launches dropped (ae_set_skip_mask — diagnostic, outputs are garbage) and timed.
Usage: python tools/kernel_share.py [--B 2 100]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audioeditingcode_b200.ddm_inversion.inversion_utils import _loop_text  # noqa: E402
from audioeditingcode_b200.unet import GraphedForward  # noqa: E402

FAMILIES = [("none", 0), ("gemm", 1), ("splitk_reduce", 2), ("gemm+reduce", 3), ("gn_stats", 4), ("gn_apply", 8),
            ("gn", 12), ("layernorm", 16), ("attention", 32), ("all", 63)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, nargs="+", default=[2, 100])
    a = ap.parse_args()
    spec = B.CONFIGS["audioldm2-large-10s"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m, cfg = B.build_model(spec, dev)
    text, cl = _loop_text(m, [""], ["a recording of a dog barking"])
    lib = m.engine.ops.lib
    for Bq in a.B:
        slot = torch.cat([torch.zeros(Bq // 2, dtype=torch.int32), torch.ones(Bq // 2, dtype=torch.int32)]).to(dev)
        base = None
        for name, mask in FAMILIES:
            lib.ae_set_skip_mask(mask)
            try:
                l0 = m.engine.ops.launch_count()
                g = GraphedForward(m.engine, Bq, spec["H"], spec["W"], text, slot, None)
                n_k = (m.engine.ops.launch_count() - l0) // 3
            finally:
                lib.ae_set_skip_mask(0)
            reps = 20 if Bq <= 4 else 3
            g.graph.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                g.graph.replay()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            if base is None:
                base = (ms, n_k)
            print(json.dumps({"B": Bq, "dropped": name, "kernels": n_k, "eval_ms": round(ms, 3),
                              "marginal_ms": round(base[0] - ms, 3), "marginal_kernels": base[1] - n_k,
                              "us_per_dropped_kernel": round(1e3 * (base[0] - ms) / max(1, base[1] - n_k), 2)}), flush=True)
            del g


if __name__ == "__main__":
    main()
