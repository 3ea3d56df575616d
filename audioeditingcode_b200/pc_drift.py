"""Unsupervised principal-direction editing — drop-in for code/pc_drift.py (same names, signatures, returns):
PromptEmbeddings (:10-13), PCStreamChoice (:16-19), expand_for_evs (:22-26), forward_directional (:29-93),
get_eigenvectors (:96-198; subspace iteration on the posterior-mean Jacobian, including the reference's
sort-before-permute behaviour, SURVEY.md Appendix D — reproduced, not repaired), apply_drift (:201-278).

Device work: U-Net evaluations through the wrapper (UNetEngine), the CFG combine + DDIM step in one kernel
(ae_ddim_step via DDIMScheduler.step); the small dense linear algebra of the iteration (norms, QR of [D, n_ev],
sort) stays in torch (cuSOLVER) — it is O(D*n_ev^2) per iteration against two U-Net evaluations.

Multi-GPU (BASELINE config 4): `get_eigenvectors(..., group=<process group>)` shards the n_ev
directions of the power iteration across the ranks — each rank runs the U-Net only on its own rows — and
sum-all-reduces the zero-padded `[n_ev, D]` posterior-mean iterate once per iteration (NCCL over NVLink on the GPU box);
normalisation, QR and sorting are then done redundantly on every rank, so every rank returns the same tensors as a
single-process run.  This is the only data-path collective of the repo.
"""
from __future__ import annotations

from enum import Enum
from typing import Dict, List, NamedTuple, Optional, Tuple

import torch


class PromptEmbeddings(NamedTuple):
    embedding_hidden_states: torch.Tensor
    embedding_class_lables: torch.Tensor
    boolean_prompt_mask: torch.Tensor


class PCStreamChoice(Enum):
    BOTH = 1
    TEXT = 2
    UNCOND = 3


def expand_for_evs(x: torch.Tensor, n_ev: int) -> torch.Tensor:
    if x is None:
        return x
    dev = x.device
    return x.repeat(n_ev, *[1] * (len(x.shape) - 1)).to(dev)


def forward_directional(ldm_stable, xt: torch.Tensor, timestep: torch.Tensor, latent: torch.Tensor,
                        uncond_emb: PromptEmbeddings, text_emb: PromptEmbeddings, cfg_tar: torch.Tensor,
                        eta: float = 1, eigvecs: torch.Tensor = 0, amount: float = 0, double_precision: bool = False,
                        mode: PCStreamChoice = PCStreamChoice.BOTH) -> torch.Tensor:
    if double_precision:
        raise NotImplementedError("double_precision is not available on the B200 path")
    with torch.no_grad():
        input = xt + amount * eigvecs * torch.sqrt(ldm_stable.model.scheduler.alphas_cumprod[int(timestep)])
    if len(xt) > 1 and \
        ((uncond_emb.boolean_prompt_mask is not None and len(uncond_emb.boolean_prompt_mask) == 1) or
         (uncond_emb.embedding_hidden_states is not None and len(uncond_emb.embedding_hidden_states) == 1)):
        n_ev = len(xt)
        uncond_emb = PromptEmbeddings(
            embedding_hidden_states=expand_for_evs(uncond_emb.embedding_hidden_states, n_ev),
            boolean_prompt_mask=expand_for_evs(uncond_emb.boolean_prompt_mask, n_ev),
            embedding_class_lables=expand_for_evs(uncond_emb.embedding_class_lables, n_ev))
        text_emb = PromptEmbeddings(
            embedding_hidden_states=expand_for_evs(text_emb.embedding_hidden_states, n_ev),
            boolean_prompt_mask=expand_for_evs(text_emb.boolean_prompt_mask, n_ev),
            embedding_class_lables=expand_for_evs(text_emb.embedding_class_lables, n_ev))
    x_u = input if mode == PCStreamChoice.BOTH or mode == PCStreamChoice.UNCOND else xt
    x_c = input if mode == PCStreamChoice.BOTH or mode == PCStreamChoice.TEXT else xt
    with torch.no_grad():
        if hasattr(ldm_stable, "cfg_pair_eval") and x_u.is_cuda:
            # one batched, graph-cached evaluation for the pair (the reference issues two: pc_drift.py:70-81)
            eps_u, eps_c = ldm_stable.cfg_pair_eval(
                x_u, x_c, timestep,
                (uncond_emb.embedding_hidden_states, uncond_emb.embedding_class_lables, uncond_emb.boolean_prompt_mask),
                (text_emb.embedding_hidden_states, text_emb.embedding_class_lables, text_emb.boolean_prompt_mask))
        else:
            eps_u = ldm_stable.unet_forward(
                x_u, timestep=timestep, encoder_hidden_states=uncond_emb.embedding_hidden_states,
                class_labels=uncond_emb.embedding_class_lables,
                encoder_attention_mask=uncond_emb.boolean_prompt_mask)[0].sample
            eps_c = ldm_stable.unet_forward(
                x_c, timestep=timestep, encoder_hidden_states=text_emb.embedding_hidden_states,
                class_labels=text_emb.embedding_class_lables,
                encoder_attention_mask=text_emb.boolean_prompt_mask)[0].sample
    noise_pred = eps_u + cfg_tar * (eps_c - eps_u)                                               # pc_drift.py:83
    res = ldm_stable.model.scheduler.step(noise_pred, timestep, input, eta=eta, variance_noise=latent)
    return res.prev_sample, res.pred_original_sample


def get_eigenvectors(ldm_stable, xt: torch.Tensor, text_emb: PromptEmbeddings, uncond_emb: PromptEmbeddings,
                     latents: torch.Tensor, mask: torch.Tensor, t: torch.Tensor, x0_pred: torch.Tensor,
                     pc_mode: PCStreamChoice = PCStreamChoice.BOTH, const: float = 1e-3, cfg_tar: float = 3,
                     iters: int = 50, double_precision: bool = False, eta: float = 1, n_ev: int = 1, group=None
                     ) -> Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor], List[torch.Tensor],
                                Dict[int, torch.Tensor], Dict[int, torch.Tensor]]:
    """`group` (extension, default None = the reference's single-process behaviour): a torch.distributed process
    group over which the n_ev directions are sharded (see module docstring)."""
    from . import parallel as _par
    rank, ws = _par.world(group) if group is not None else (0, 1)
    rows = _par.shard_indices(n_ev, rank, ws) if ws > 1 else None
    if n_ev > 1:
        x0_pred = expand_for_evs(x0_pred, n_ev)
        xt = expand_for_evs(xt, n_ev)
        uncond_emb = PromptEmbeddings(
            embedding_hidden_states=expand_for_evs(uncond_emb.embedding_hidden_states, n_ev),
            boolean_prompt_mask=expand_for_evs(uncond_emb.boolean_prompt_mask, n_ev),
            embedding_class_lables=expand_for_evs(uncond_emb.embedding_class_lables, n_ev))
        text_emb = PromptEmbeddings(
            embedding_hidden_states=expand_for_evs(text_emb.embedding_hidden_states, n_ev),
            boolean_prompt_mask=expand_for_evs(text_emb.boolean_prompt_mask, n_ev),
            embedding_class_lables=expand_for_evs(text_emb.embedding_class_lables, n_ev))
    eigvecs = torch.randn_like(xt) * mask * const
    if rows is not None:
        _par.broadcast_(eigvecs, 0, group)      # one random start for all ranks
    prev_ev = eigvecs.detach().clone()
    in_corr, in_norm = [], []
    interm_eigvecs, interm_eigvals = {}, {}
    with torch.no_grad():
        for i in range(iters):
            if rows is None:
                _, unmaksed_out = forward_directional(ldm_stable, xt, t, latents, uncond_emb, text_emb, cfg_tar,
                                                      eta=eta, eigvecs=eigvecs, amount=1,
                                                      double_precision=double_precision, mode=pc_mode)
            else:
                local = None
                if rows:
                    pick = (lambda v: v if v is None or len(v) != n_ev else v[rows])
                    _, local = forward_directional(
                        ldm_stable, xt[rows], t, pick(latents), PromptEmbeddings(*[pick(v) for v in uncond_emb]),
                        PromptEmbeddings(*[pick(v) for v in text_emb]), cfg_tar, eta=eta, eigvecs=eigvecs[rows],
                        amount=1, double_precision=double_precision, mode=pc_mode)
                unmaksed_out = _par.allreduce_rows(local, rows, xt, group)
            out = unmaksed_out * mask
            Ab = out - x0_pred
            if n_ev > 1:
                if len(xt.shape) == 4:
                    permute_arg = (1, 2, 3, 0)
                elif len(xt.shape) == 3:
                    permute_arg = (1, 2, 0)
                elif len(xt.shape) == 2:
                    permute_arg = (1, 0)
                norm_of_Ab = Ab[:, mask[0].to(torch.bool)].norm(dim=1)
                eigvecs = (Ab / norm_of_Ab.reshape(n_ev, *[1] * (len(xt.shape) - 1))) * mask
                Q, R = torch.linalg.qr(eigvecs.permute(*permute_arg).reshape(-1, n_ev), mode='reduced')
                swap = torch.prod(torch.linalg.diagonal(R))
                if swap < 0:
                    Q *= -1
                eigvecs = Q / Q.norm(dim=0)
                eigvecs = eigvecs.T.reshape(Ab.shape)
                _, tmp = (norm_of_Ab / const * (ldm_stable.get_sigma(t) ** 2)).reshape(n_ev, ).sort(
                    descending=True, stable=True)
                eigvecs = eigvecs[tmp, ...]
            else:
                norm_of_Ab = Ab[mask.to(torch.bool)].norm()
                eigvecs = (Ab / norm_of_Ab) * mask
            if i > 0:
                corr = ((prev_ev.reshape(n_ev, -1)) @ (eigvecs.reshape(n_ev, -1).T)).diag()
                in_corr.append(corr)
            in_norm.append(norm_of_Ab)
            prev_ev = eigvecs.detach().clone()
            if not (i % 10) and i > 15:
                interm_eigvecs[i] = eigvecs
                interm_eigvals[i] = norm_of_Ab / const * (ldm_stable.get_sigma(t) ** 2)
            eigvecs *= const
    eigval = (norm_of_Ab / const * (ldm_stable.get_sigma(t) ** 2))
    eigvecs /= const
    return eigvecs, eigval, in_corr, in_norm, interm_eigvecs, interm_eigvals


def apply_drift(ldm_stable, xt_m1: torch.Tensor, x0_pred: torch.Tensor, t: torch.Tensor, timesteps: torch.Tensor,
                num_diff_steps: int, eigdata: Dict[int, Dict[str, torch.Tensor]], latent: torch.Tensor,
                device: torch.device, use_shifted_x0_for_noisepred: bool = True,
                use_specific_ts_pc: Optional[int] = None, amount: float = 1, sub_iters: Optional[int] = None,
                eta: float = 1, ev_nums: List[int] = [1], evals: Optional[Dict[int, torch.Tensor]] = None
                ) -> torch.Tensor:
    if use_specific_ts_pc is None:
        use_t = t.item()
    else:
        use_t = timesteps[num_diff_steps - use_specific_ts_pc].item()
    eigvec = eigdata[use_t]['eigvec'].to(device)
    if evals is None:
        eigval = eigdata[t.item()]['eigval'].to(device)
    else:
        eigval = torch.from_numpy(evals[t.item()]).to(device)
    if sub_iters is not None:
        eigvec = eigdata[use_t]['interm_eigvecs'][sub_iters].to(device)
        if evals is not None:
            raise ValueError("evals should be None if sub_iters is not None")
        eigval = eigdata[t.item()]['interm_eigvals'][sub_iters].to(device)
    shift_by = 0
    for ev_num in ev_nums:
        ev_idx = ev_num - 1
        shift_by += amount * (eigval[ev_idx].unsqueeze(0).sqrt() * eigvec[ev_idx].unsqueeze(0))
    x0_pred_drift = x0_pred.clone() + shift_by
    sched = ldm_stable.model.scheduler
    prev_timestep = t - sched.config.num_train_timesteps // sched.num_inference_steps
    variance = sched._get_variance(t, prev_timestep)
    std_dev_t = eta * variance ** (0.5)
    alpha_prod_t_prev = sched.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else sched.final_alpha_cumprod
    alpha_prod_t = sched.alphas_cumprod[int(t)]
    beta_prod_t = 1 - alpha_prod_t
    if eta > 0:
        xt_m1 = xt_m1 - std_dev_t * latent
    pred_sample_direction = xt_m1 - alpha_prod_t_prev ** (0.5) * x0_pred
    pred_epsilon = pred_sample_direction / ((1 - alpha_prod_t_prev - std_dev_t ** 2) ** (0.5))
    if use_shifted_x0_for_noisepred:
        pred_epsilon = pred_epsilon - (alpha_prod_t ** (0.5)) / (beta_prod_t ** (0.5)) * shift_by
    pred_sample_direction = (1 - alpha_prod_t_prev - std_dev_t ** 2) ** (0.5) * pred_epsilon
    xt_m1 = alpha_prod_t_prev ** (0.5) * x0_pred_drift + pred_sample_direction
    if eta > 0:
        xt_m1 = xt_m1 + std_dev_t * latent
    return xt_m1
