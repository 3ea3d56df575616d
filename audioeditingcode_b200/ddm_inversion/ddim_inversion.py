"""DDIM inversion baseline — drop-in for code/ddm_inversion/ddim_inversion.py (`--mode ddim` of main_run.py):
next_step (:10-20), get_noise_pred (:23-41), ddim_inversion (:44-56), text2image_ldm_stable (:59-84)."""
from __future__ import annotations

from typing import List, Optional, Union

import torch
from tqdm import tqdm

from ..models import PipelineWrapper


def next_step(ldm_model: PipelineWrapper, model_output: torch.Tensor, timestep: int, sample: torch.Tensor) -> torch.Tensor:
    sched = ldm_model.model.scheduler
    timestep, next_timestep = min(timestep - sched.config.num_train_timesteps // sched.num_inference_steps, 999), timestep
    alpha_prod_t = sched.alphas_cumprod[int(timestep)] if timestep >= 0 else sched.final_alpha_cumprod
    alpha_prod_t_next = sched.alphas_cumprod[int(next_timestep)]
    beta_prod_t = 1 - alpha_prod_t
    next_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
    next_sample_direction = (1 - alpha_prod_t_next) ** 0.5 * model_output
    return alpha_prod_t_next ** 0.5 * next_original_sample + next_sample_direction


def get_noise_pred(ldm_model: PipelineWrapper, latent: torch.Tensor, t: torch.Tensor, text_emb, uncond_emb,
                   cfg_scale: float) -> torch.Tensor:
    text_hs, text_cl, text_mask = text_emb
    un_hs, un_cl, un_mask = uncond_emb
    with torch.no_grad():
        if hasattr(ldm_model, "cfg_pair_eval") and latent.is_cuda:
            # one batched, graph-cached evaluation for the pair (the reference issues two: ddim_inversion.py:31-38)
            eps_u, eps_c = ldm_model.cfg_pair_eval(latent, latent, t, (un_hs, un_cl, un_mask), (text_hs, text_cl, text_mask))
            return eps_u + cfg_scale * (eps_c - eps_u)
        uncond_out, _, _ = ldm_model.unet_forward(latent, timestep=t, encoder_hidden_states=un_hs, class_labels=un_cl,
                                                  encoder_attention_mask=un_mask)
        cond_out, _, _ = ldm_model.unet_forward(latent, timestep=t, encoder_hidden_states=text_hs, class_labels=text_cl,
                                                encoder_attention_mask=text_mask)
    return uncond_out.sample + cfg_scale * (cond_out.sample - uncond_out.sample)


def ddim_inversion(ldm_model: PipelineWrapper, w0: torch.Tensor, prompts: List[str], cfg_scale: float,
                   num_inference_steps: int, skip: Union[int, torch.Tensor] = 0) -> torch.Tensor:
    text_emb = ldm_model.encode_text(prompts)
    uncond_emb = ldm_model.encode_text([""])
    latent = w0.clone().detach()
    ts = ldm_model.model.scheduler.timesteps_cpu
    for i in tqdm(range(num_inference_steps)):
        if num_inference_steps - i <= skip:
            break
        t = ts[len(ts) - i - 1]
        noise_pred = get_noise_pred(ldm_model, latent, t, text_emb, uncond_emb, cfg_scale)
        latent = next_step(ldm_model, noise_pred, int(t), latent)
    return latent


@torch.no_grad()
def text2image_ldm_stable(ldm_model: PipelineWrapper, prompt: List[str], num_inference_steps: int = 50,
                          guidance_scale: float = 7.5, xt: Optional[torch.Tensor] = None,
                          skip: Union[int, torch.Tensor] = 0) -> torch.Tensor:
    text_emb = ldm_model.encode_text(prompt)
    uncond_emb = ldm_model.encode_text([""])
    for t in tqdm(ldm_model.model.scheduler.timesteps_cpu[int(skip):]):
        noise_pred = get_noise_pred(ldm_model, xt, t, text_emb, uncond_emb, guidance_scale)
        xt = ldm_model.model.scheduler.step(noise_pred, t, xt, eta=0).prev_sample
    return xt
