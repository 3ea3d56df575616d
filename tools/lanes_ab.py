"""A/B of the two-lane execution (forward chunks || reverse steps) on the bench workload, one model build.
Usage: python tools/lanes_ab.py [--fb 50 25] — prints ms per 300-step job for overlap off / on per forward batch."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audioeditingcode_b200.ddm_inversion import inversion_utils as IU  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fb", type=int, nargs="+", default=[50, 25])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--config", default="audioldm2-large-10s")
    a = ap.parse_args()
    spec = B.CONFIGS[a.config]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m, cfg = B.build_model(spec, dev)
    x0 = (0.5 * torch.randn(1, cfg.in_channels, spec["H"], spec["W"], generator=torch.Generator().manual_seed(1))).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    for fb in a.fb:
        for mode in (False, True):
            IU.OVERLAP = mode
            for _ in range(3):
                w = B.run_job(m, spec, x0, fb)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(a.reps):
                flush.zero_()
                w = B.run_job(m, spec, x0, fb)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / a.reps
            steps = spec["n_inv"] + spec["tstart"]
            same = None
            if fb == a.fb[0]:
                if ref is None:
                    ref = w.clone()
                else:
                    same = bool(torch.equal(ref, w))
            print(json.dumps({"forward_batch": fb, "overlap": mode, "ms_per_job": round(ms, 2),
                              "steps_per_s": round(steps / ms * 1e3, 1), "overlap_hits": getattr(m, "overlap_hits", 0),
                              "same_bits_as_first": same}), flush=True)


if __name__ == "__main__":
    main()
