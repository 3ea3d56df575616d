"""Shared test helpers (test infrastructure: may import oracle/)."""
import os

import numpy as np
import torch

from oracle import unet_torch as U
from oracle.ddpm_oracle import MiniDDIM
from audioeditingcode_b200 import unet_config as C

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: (torch.from_numpy(d[k]) if d[k].ndim > 0 else d[k].item()) for k in d.files}


def tiny_cfg_and_weights(name="tiny-audioldm", seed=0):
    cfg = C.preset(name)
    return cfg, U.synthetic_weights(cfg, seed=seed)


def oracle_unet_fn(cfg, w, uncond_y, cond_y):
    """UNetFn for oracle.ddpm_oracle loops using the torch restatement with AudioLDM class labels."""
    def fn(x, t, which):
        y = uncond_y if which == "uncond" else cond_y
        tt = torch.full((x.shape[0],), int(t), dtype=torch.int64)
        with torch.no_grad():
            return U.unet_forward(cfg, w, x, tt, class_labels=y)[0]
    return fn


def make_sched(cfg, n, pred=None):
    s = MiniDDIM(cfg.beta_start, cfg.beta_end, prediction_type=pred or cfg.prediction_type)
    s.set_timesteps(n)
    return s
