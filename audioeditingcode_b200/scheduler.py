"""DDIM scheduler object with the surface the reference's loops touch on `model.model.scheduler`
(SURVEY.md §8b): timesteps, num_inference_steps, alphas_cumprod (CPU fp32, indexed on the host exactly like
code/models.py:89-91), final_alpha_cumprod, config.{num_train_timesteps,prediction_type}, scale_model_input,
init_noise_sigma, add_noise, step, _get_variance.

[UPSTREAM] restatement of diffusers.DDIMScheduler as configured by the checkpoints the reference loads
(beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=1, set_alpha_to_one=False,
clip_sample=False).  Integer index math is exact; the device arithmetic of `step`
runs in libaedit (ae_ddim_step).
"""
from __future__ import annotations

import ctypes as C
import types
from typing import Optional

import torch

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SchedTable:
    """libaedit scheduler table bound to a DDIMScheduler state (rebuilt when set_timesteps changes).

    The per-step scalars are computed HERE with the reference's own expressions on 0-dim fp32 CPU tensors
    (code/models.py:89-111,126-150, :539-549) and handed to the library (ae_sched_create_from_rows /
    ae_sched_set_eta): torch's CPU `**0.5` is not always the correctly rounded square root, and the reference
    evaluates these scalars on the host even when the latents live on the GPU (alphas_cumprod stays on the CPU)."""

    def __init__(self, scheduler: "DDIMScheduler"):
        lib = _lib.load()
        self.lib = lib
        ac = scheduler.alphas_cumprod
        ts = scheduler.timesteps_cpu
        N = ts.numel()
        step = scheduler.config.num_train_timesteps // N                       # models.py:96-97
        rows = (_lib.AeSchedRow * N)()
        self._ap, self._var = [], []
        for k in range(N):
            t = int(ts[k])
            prev = t - step
            ab = ac[t]
            ap = ac[prev] if prev >= 0 else scheduler.final_alpha_cumprod      # models.py:547-549
            var = ((1 - ap) / (1 - ab)) * (1 - ab / ap)                        # models.py:539-545
            r = rows[k]
            r.t, r.prev_t = t, prev
            r.alpha_bar_t, r.alpha_prod_t_prev, r.variance = float(ab), float(ap), float(var)
            r.sqrt_ab = float(ab ** 0.5)
            r.sqrt_1mab = float((1 - ab) ** 0.5)
            r.sqrt_ap = float(ap ** 0.5)
            r.sqrt_var = float(var ** 0.5)
            self._ap.append(ap)
            self._var.append(var)
        h = C.c_void_p()
        pred = {"epsilon": 0, "v_prediction": 1}[scheduler.config.prediction_type]
        _lib.check(lib.ae_sched_create_from_rows(rows, N, pred, scheduler.config.num_train_timesteps, C.byref(h)),
                   "ae_sched_create_from_rows")
        self.h = h
        self.N = N
        self._eta_key = None

    def set_etas(self, etas) -> None:
        """etas: list of length N indexed like the reference's `etas[idx]`, idx = N - pos - 1."""
        key = tuple(float(e) for e in etas)
        if key == self._eta_key:
            return
        if self._eta_key is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
            torch.cuda.synchronize()      # kernels of a run in flight on a side lane may still read the old table
        N = self.N
        cdir = (C.c_float * N)()
        sig = (C.c_float * N)()
        for pos in range(N):
            eta = key[N - pos - 1]
            ap, var = self._ap[pos], self._var[pos]
            cdir[pos] = float((1 - ap - eta * var) ** 0.5)                      # models.py:107 / :148
            sig[pos] = float(eta * var ** 0.5)                                   # models.py:111 / :155
        _lib.check(self.lib.ae_sched_set_eta(self.h, C.cast(cdir, C.c_void_p), C.cast(sig, C.c_void_p)),
                   "ae_sched_set_eta")
        self._eta_key = key

    def pos_of_t(self, t: int) -> int:
        pos = self.lib.ae_sched_pos_of_t(self.h, int(t))
        if pos < 0:
            raise KeyError(f"timestep {int(t)} is not in the scheduler's timesteps")   # dict KeyError in the reference
        return pos

    def row(self, pos: int) -> _lib.AeSchedRow:
        r = _lib.AeSchedRow()
        _lib.check(self.lib.ae_sched_row_h(self.h, pos, C.byref(r)), "ae_sched_row_h")
        return r

    def __del__(self):
        try:
            self.lib.ae_sched_destroy(self.h)
        except Exception:
            pass




class DDIMScheduler:
    def __init__(self, beta_start: float = 0.0015, beta_end: float = 0.0195, num_train_timesteps: int = 1000,
                 prediction_type: str = "epsilon", steps_offset: int = 1):
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.config = types.SimpleNamespace(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                                            steps_offset=steps_offset, beta_start=beta_start, beta_end=beta_end)
        self.num_inference_steps: Optional[int] = None
        self.timesteps: Optional[torch.Tensor] = None
        self._table = None   # libaedit scheduler table, built lazily by the wrapper

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError("num_inference_steps larger than num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps, dtype=torch.int64) * step_ratio).flip(0) + self.config.steps_offset
        self.timesteps_cpu = ts
        self.timesteps = ts.to(device) if device is not None else ts
        self._table = None

    @property
    def table(self) -> SchedTable:
        if self._table is None:
            self._table = SchedTable(self)
        return self._table

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        """[UPSTREAM] DDIMScheduler.step (std = eta*sqrt(var); direction uses std**2) on the device (ae_ddim_step).
        Returns an object with .prev_sample and .pred_original_sample (what pc_drift.py:89-93 reads)."""
        tab = self.table
        pos = tab.pos_of_t(int(timestep))
        mo = model_output.to(torch.float32).contiguous()
        x = sample.to(torch.float32).contiguous()
        if eta > 0 and variance_noise is None:
            variance_noise = torch.randn(mo.shape, generator=generator, device=mo.device, dtype=mo.dtype)
        vn = None if variance_noise is None else variance_noise.to(torch.float32).expand_as(mo).contiguous()
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x)
        _lib.check(tab.lib.ae_ddim_step(tab.h, pos, float(eta), 0.0, _ptr(mo), None, _ptr(x), _ptr(vn), _ptr(prev),
                                        _ptr(x0), x.numel(), _stream()), "ae_ddim_step")
        return types.SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[int(timestep)]
        a_p = self.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        so = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        so = so.flatten()
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            so = so.unsqueeze(-1)
        return sa * original_samples + so * noise
