"""Per-block error attribution of one full-size U-Net evaluation: CUDA engine (eager, probes on) vs the fp32 CPU
oracle (oracle/unet_torch.py probes of the same names).  Prints the cumulative rel-L2 error after conv_in and after
every ResBlock / attention site / resampler, and of the final eps — to locate which layers carry the operand-precision
error ("nobody has located which layers carry the error").

    python tools/error_attribution.py [audioldm2-large|tango|audioldm-s] [t]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import unet_torch as U                      # noqa: E402
from audioeditingcode_b200 import unet_config as C      # noqa: E402
from tests import fullsize as FS                        # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "audioldm2-large"
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 501
    cfg = C.preset(name)
    w = U.synthetic_weights(cfg, seed=0)
    from audioeditingcode_b200.unet import UNetEngine
    eng = UNetEngine(cfg, w, "cuda")
    streams, masks, cl = FS.text_rows(cfg, 2)
    gen = torch.Generator().manual_seed(7)
    x = 0.8 * torch.randn(1, 8, 256, 16, generator=gen).expand(2, -1, -1, -1).contiguous()
    ref_p, got_p = {}, {}
    ref = FS.oracle_eval(cfg, w, x, t, streams, masks, cl, rows=[0, 1], probe=lambda n, v: ref_p.__setitem__(n, v.clone()))

    def probe(n, v, hw):
        B = v.shape[0]
        got_p[n] = v.detach().float().reshape(B, hw[0], hw[1], -1).permute(0, 3, 1, 2).cpu()
    eng.probe = probe
    kw = {}
    if streams:
        kw["text"] = eng.prepare_text([s.cuda() for s in streams], [None if m is None else m.cuda() for m in masks])
        kw["slot_map"] = torch.tensor([0, 1], dtype=torch.int32).cuda()
    if cl is not None:
        kw["class_labels"] = cl.cuda()
    out = eng.forward(x.cuda(), torch.full((2,), t, dtype=torch.int64).cuda(), **kw).cpu()
    print(f"# {name} t={t} B=2 latent [2,8,256,16]; operand dtype {eng.adt}")
    print(f"{'block':42s} {'rel-L2':>10s} {'max-abs':>10s} {'|ref|_rms':>10s}")
    for n in ref_p:
        a, b = got_p[n], ref_p[n]
        print(f"{n:42s} {((a - b).norm() / b.norm()).item():10.3e} {(a - b).abs().max().item():10.3e} "
              f"{b.pow(2).mean().sqrt().item():10.3e}")
    print(f"{'eps (output)':42s} {((out - ref).norm() / ref.norm()).item():10.3e} {(out - ref).abs().max().item():10.3e} "
          f"{ref.pow(2).mean().sqrt().item():10.3e}")


if __name__ == "__main__":
    main()
