"""Time one captured U-Net evaluation (B=2 reverse shape, B=100 forward chunk) under library settings.
Usage: python tools/eval_time.py --pdlx 0 3 4 7 8 15"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audioeditingcode_b200.ddm_inversion.inversion_utils import _loop_text  # noqa: E402
from audioeditingcode_b200.unet import GraphedForward  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, nargs="+", default=[2, 100])
    ap.add_argument("--pdlx", type=int, nargs="+", default=[0])
    ap.add_argument("--env", nargs="*", default=[], help="lib setter calls name=int, e.g. ae_set_tile_model=0")
    a = ap.parse_args()
    spec = B.CONFIGS["audioldm2-large-10s"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m, cfg = B.build_model(spec, dev)
    text, cl = _loop_text(m, [""], ["a recording of a dog barking"])
    lib = m.engine.ops.lib
    for kv in a.env:
        k, v = kv.split("=")
        getattr(lib, k)(*[int(x) for x in v.split(",")])
    for Bq in a.B:
        slot = torch.cat([torch.zeros(Bq // 2, dtype=torch.int32), torch.ones(Bq // 2, dtype=torch.int32)]).to(dev)
        for px in a.pdlx:
            lib.ae_set_pdl_extra(px)
            g = GraphedForward(m.engine, Bq, spec["H"], spec["W"], text, slot, None)
            reps = 30 if Bq <= 4 else 4
            best = 1e9
            for _ in range(3):
                g.graph.replay()
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(reps):
                    g.graph.replay()
                e.record()
                torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e) / reps)
            print(json.dumps({"B": Bq, "pdl_extra": px, "eval_ms": round(best, 3)}), flush=True)
            del g
    lib.ae_set_pdl_extra(0)


if __name__ == "__main__":
    main()
