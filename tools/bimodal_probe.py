"""A B=2 evaluation graph lands on ~5.97 or ~6.20 ms depending on the capture: probe what differs (addresses)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audioeditingcode_b200.ddm_inversion.inversion_utils import _loop_text  # noqa: E402
from audioeditingcode_b200.unet import GraphedForward  # noqa: E402

spec = B.CONFIGS["audioldm2-large-10s"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
m, cfg = B.build_model(spec, dev)
text, cl = _loop_text(m, [""], ["a recording of a dog barking"])
slot = torch.tensor([0, 1], dtype=torch.int32, device=dev)
keep = []
for i in range(10):
    g = GraphedForward(m.engine, 2, spec["H"], spec["W"], text, slot, None, lane=(1 if i % 2 else 0))
    best = 1e9
    for _ in range(3):
        g.graph.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(30):
            g.graph.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 30)
    arena = m.engine._cs_arena
    print(json.dumps({"i": i, "lane": i % 2, "eval_ms": round(best, 3), "out_ptr": hex(g.out.data_ptr()),
                      "x_ptr": hex(g.x.data_ptr()), "arena_ptr": hex(arena.data_ptr()) if arena is not None else None,
                      "out_mod_2MB": g.out.data_ptr() % (2 << 20), "arena_mod_2MB": arena.data_ptr() % (2 << 20)}), flush=True)
    if i % 3 == 0:
        keep.append(g)       # vary the allocator state
    else:
        del g
