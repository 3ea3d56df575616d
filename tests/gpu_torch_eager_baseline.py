"""Informational GPU baseline ("also time stock PyTorch-CUDA eager ... on the same B200").

Runs the ORACLE restatement of the U-Net (oracle/unet_torch.py: plain torch.nn.functional ops -> cuDNN / cuBLAS /
SDPA kernels) on cuda:0 in the reference's calling pattern — two separate B=1 U-Net evaluations per CFG step
(inversion_utils.py:86-101, 249-276), scheduler math in torch — for a bounded number of CFG steps of the bench
workload, in fp32 (TF32 off), fp32 with TF32 allowed, and bf16 autocast.  This is test/measurement infrastructure
(it imports oracle/), never part of the product path and never the bench value.

    python -m tests.gpu_torch_eager_baseline [--config audioldm2-large-10s] [--steps 20]
prints one JSON line per precision mode.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def run(spec, mode, n_steps, batched, device="cuda:0"):
    from oracle import unet_torch as U
    from oracle import ddpm_oracle as D
    from audioeditingcode_b200 import unet_config as C
    dev = torch.device(device)
    cfg = C.preset(spec["preset"])
    w = {k: v.to(dev) for k, v in U.synthetic_weights(cfg, seed=0).items()}
    g = torch.Generator().manual_seed(1)
    H, Wd = spec["H"], spec["W"]
    x0 = (0.5 * torch.randn(1, cfg.in_channels, H, Wd, generator=g)).to(dev)
    dims = {s[1]: s[0] for s in cfg.transformer_specs if s is not None}
    lens = spec["text_lens"]
    su = [torch.randn(1, lens[i], dims[i], generator=g).to(dev) for i in range(cfg.n_streams)]
    sc = [torch.randn(1, lens[i], dims[i], generator=g).to(dev) for i in range(cfg.n_streams)]
    yu = yc = None
    if cfg.class_embed_dim:
        yu = torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1).to(dev)
        yc = torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1).to(dev)
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    ac = torch.autocast(dev.type if dev.type != "meta" else "cpu", dtype=torch.bfloat16, enabled=(mode == "bf16"))

    def unet(x, t, st, y):
        tt = torch.full((x.shape[0],), int(t), dtype=torch.int64, device=dev)
        with torch.no_grad(), ac:
            return U.unet_forward(cfg, w, x, tt, streams=st, stream_masks=[None] * len(st), class_labels=y)[0].float()

    sched = D.MiniDDIM(cfg.beta_start, cfg.beta_end, prediction_type=cfg.prediction_type)
    sched.set_timesteps(spec["n_inv"])
    N = spec["n_inv"]
    noise = torch.randn(N, *x0.shape[1:], generator=g).to(dev)
    xts = D.sample_xts_from_x0(sched, x0, noise)
    cfgm, _ = D.build_cfg_maps(1, x0.shape[1:], [spec["cfg_src"]], None)
    cfgm = cfgm.to(dev)

    def one_step(pos):
        t = int(sched.timesteps[pos])
        idx = N - pos - 1
        xt = xts[idx + 1][None]
        if batched:
            both = unet(torch.cat([xt, xt]), t, [torch.cat([a, b]) for a, b in zip(su, sc)],
                        None if yu is None else torch.cat([yu, yc]))
            eu, ec = both[:1], both[1:]
        else:
            eu, ec = unet(xt, t, su, yu), unet(xt, t, sc, yc)
        eps = D.cfg_combine(eu, ec, cfgm)
        if pos % 2 == 0:
            D.get_zs_from_xts(sched, xt, xts[idx][None], eps, t, 1.0, True)
        else:
            D.reverse_step_with_custom_noise(sched, eps, t, xt, noise[idx][None], 1.0)
    for k in range(3):
        one_step(k)
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    sync()
    t0 = time.perf_counter()
    for k in range(n_steps):
        one_step(3 + k)
    sync()
    dt = time.perf_counter() - t0
    return n_steps / dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="audioldm2-large-10s", choices=sorted(bench.CONFIGS))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args()
    spec = bench.CONFIGS[args.config]
    for mode in ("fp32", "tf32", "bf16"):
        for batched in (False, True):
            try:
                v = run(spec, mode, args.steps, batched, args.device)
                print(json.dumps({"impl": "torch-cuda-eager(oracle restatement)", "mode": mode,
                                  "cfg_pair": "one B=2 call" if batched else "two B=1 calls (reference pattern)",
                                  "metric": "denoising-steps/sec", "value": v, "steps_timed": args.steps,
                                  "config": args.config}), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"impl": "torch-cuda-eager", "mode": mode, "batched": batched, "error": repr(e)[:300]}),
                      flush=True)


if __name__ == "__main__":
    main()
