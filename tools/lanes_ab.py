"""A/B of execution variants on the bench workload with one model build: ms per 300-step job.
Variant syntax: key=value pairs joined by ',' — overlap(0/1) fb head rev(adaptive/solo/shared) pdlx gnstream(0/1)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audioeditingcode_b200.ddm_inversion import inversion_utils as IU  # noqa: E402

DEFAULTS = dict(overlap=1, fb=50, head=10, rev="adaptive", pdlx=0, gnstream=1, pf=0, ps=0, pc=15, skb=0, pmt=296, gpt=4, rep=0, ssm=1)


def timed(fn, reps, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        flush.zero_()
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--config", default="audioldm2-large-10s")
    ap.add_argument("--variants", nargs="+", default=["overlap=0", "head=0", "head=10", "head=5", "head=20"])
    a = ap.parse_args()
    spec = B.CONFIGS[a.config]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m, cfg = B.build_model(spec, dev)
    lib = m.engine.ops.lib
    steps = spec["n_inv"] + spec["tstart"]
    x0 = (0.5 * torch.randn(1, cfg.in_channels, spec["H"], spec["W"], generator=torch.Generator().manual_seed(1))).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    last = None
    for v in a.variants:
        o = dict(DEFAULTS)
        for kv in v.split(","):
            k, val = kv.split("=")
            o[k] = val if k == "rev" else int(val)
        key = (o["pdlx"], o["gnstream"], o["pf"], o["ps"], o["pc"], o["skb"], o["pmt"], o["gpt"], o["rep"], o["ssm"])
        if key != last:
            torch.cuda.synchronize()
            m.engine._graphs.clear()
            m.engine.pdl_extra = [o["pf"] or o["pdlx"], o["ps"] or o["pdlx"], o["pc"] or o["pdlx"]]
            lib.ae_set_shallow_kblocks(o["skb"])
            lib.ae_set_persistent_min_tiles(o["pmt"])
            m.engine.graph_placement_tries = o["gpt"]
            m.engine.shared_sm_rings = bool(o["ssm"])
            lib.ae_set_gn_stream_min_bytes((8 << 20) if o["gnstream"] else (1 << 60))
            last = key
        IU.OVERLAP = bool(o["overlap"])
        IU.REV_VARIANT = o["rev"]
        IU.DEFAULT_HEAD_CHUNK = o["head"]
        h0 = getattr(m, "overlap_hits", 0)
        ms = timed(lambda: B.run_job(m, spec, x0, o["fb"]), a.reps, flush)
        place = {str(k[0]) + "/lane" + str(k[-1]): getattr(g, "placement_ms", None) for k, g in m.engine._graphs.items()
                 if k[0] <= 4}
        print(json.dumps({"variant": v, "ms_per_job": round(ms, 2), "steps_per_s": round(steps / ms * 1e3, 1),
                          "overlap_hits": getattr(m, "overlap_hits", 0) - h0, "placement_ms": place}), flush=True)


if __name__ == "__main__":
    main()
