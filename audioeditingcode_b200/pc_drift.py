"""Unsupervised principal-direction editing on the B200 path — drop-in for code/pc_drift.py (same public names,
signatures and return tuples): PromptEmbeddings (:10-13), PCStreamChoice (:16-19), expand_for_evs (:22-26),
forward_directional (:29-93), get_eigenvectors (:96-198), apply_drift (:201-278).

How it runs here (every tensor op of the iteration is a libaedit kernel, csrc/pc_kernels.cu):

  forward_directional   ae_pc_perturb writes `xt + amount*eigvecs*sqrt(alpha_bar_t)` straight into the 2n-row CFG batch
                        (uncond rows | cond rows, per PCStreamChoice) -> ONE graph-cached U-Net evaluation for the pair
                        (the reference issues two, :64-80) -> ae_ddim_step = CFG combine (:83) + DDIMScheduler.step (:89)
  get_eigenvectors      per subspace iteration: forward_directional on the n_ev perturbed rows, then
                        ae_pc_subspace_step = masked difference, per-direction norms, normalisation, re-orthonormalisation
                        (CholeskyQR2 with LAPACK's Householder sign rule and the reference's `swap` / sort-before-permute
                        behaviour, SURVEY.md Appendix D — reproduced, not repaired), correlation with the previous iterate
                        and the next perturbation `const * eigvecs`.  No host synchronisation inside the loop (the
                        reference syncs on `if swap < 0` every iteration, :165).
  apply_drift           ae_pc_apply_drift: one elementwise kernel for :232-278

Multi-GPU (BASELINE configs[3]): `get_eigenvectors(..., group=<process group>)` shards the n_ev
directions over the ranks — each rank runs the U-Net only on its own rows — and ALL-GATHERS the owned rows of the
posterior-mean iterate once per iteration (NCCL over NVLink: 1/world of the bytes of a zero-padded all-reduce); the
n x n algebra then runs redundantly, so every rank returns the same tensors as a single-process run.
"""
from __future__ import annotations

import ctypes as C
from enum import Enum
from typing import Dict, List, NamedTuple, Optional, Tuple

import torch

from . import _lib


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
    return x.repeat(n_ev, *[1] * (len(x.shape) - 1)).to(x.device)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


def _triple(e: PromptEmbeddings):
    return (e.embedding_hidden_states, e.embedding_class_lables, e.boolean_prompt_mask)


def forward_directional(ldm_stable, xt: torch.Tensor, timestep: torch.Tensor, latent: torch.Tensor,
                        uncond_emb: PromptEmbeddings, text_emb: PromptEmbeddings, cfg_tar: torch.Tensor,
                        eta: float = 1, eigvecs: torch.Tensor = 0, amount: float = 0, double_precision: bool = False,
                        mode: PCStreamChoice = PCStreamChoice.BOTH) -> Tuple[torch.Tensor, torch.Tensor]:
    """One CFG denoising step, optionally with the input shifted along `eigvecs` (pc_drift.py:29-93).  Returns
    (prev_sample, pred_original_sample).  Embeddings with one row are shared by all rows of `xt` (the reference
    repeats them, :46-58; here the text K/V of the single row are indexed by every sample)."""
    if double_precision:
        raise NotImplementedError("double_precision is not available on the B200 path")
    lib = _lib.load()
    dev = ldm_stable.device
    sched = ldm_stable.model.scheduler
    xt = _f32(xt, dev)
    n = xt.shape[0]
    n_el = xt[0].numel()
    ev = None
    if torch.is_tensor(eigvecs) and amount != 0:
        ev = _f32(eigvecs, dev)
        if ev.shape[0] != n:
            ev = ev.expand(n, *ev.shape[1:]).contiguous()
    # torch.sqrt on the host scalar like the reference (:42); alphas_cumprod lives on the CPU
    sqrt_ab = float(torch.sqrt(sched.alphas_cumprod[int(timestep)]))
    x_batch = torch.empty((2 * n, *xt.shape[1:]), device=dev, dtype=torch.float32)
    inp = torch.empty_like(xt)
    _lib.check(lib.ae_pc_perturb(_p(xt), n_el, _p(ev), float(amount), sqrt_ab, int(mode.value), n, _p(x_batch), _p(inp),
                                 n_el, _stream()), "ae_pc_perturb")
    eps = ldm_stable.cfg_pair_eval_batch(x_batch, timestep, _triple(uncond_emb), _triple(text_emb))
    tab = sched.table
    pos = tab.pos_of_t(int(timestep))
    vn = None
    if eta > 0:
        if latent is None:
            latent = torch.randn(xt.shape, device=dev)
        vn = _f32(latent, dev)
        if vn.shape[0] != n:
            vn = vn.expand(n, *vn.shape[1:]).contiguous()
    prev = torch.empty_like(xt)
    x0p = torch.empty_like(xt)
    _lib.check(lib.ae_ddim_step(tab.h, pos, float(eta), float(cfg_tar), _p(eps), _p(eps[n:]), _p(inp), _p(vn), _p(prev),
                                _p(x0p), xt.numel(), _stream()), "ae_ddim_step")
    return prev, x0p


def resolve_fd_const(ldm_stable, const: float, fd_const: Optional[float] = None) -> float:
    """The finite-difference step get_eigenvectors actually takes (see its docstring): the `fd_const` argument, else env
    AEDIT_PC_FD_CONST (a number, or "reference" / "const" for the caller's `const`), else the evaluator's own
    `pc_fd_const` attribute, else `const` — the reference's behaviour (pc_drift.py:130,140)."""
    import os as _os
    if fd_const is None:
        env = _os.environ.get("AEDIT_PC_FD_CONST", "")
        if env:
            fd_const = None if env.lower() in ("reference", "const") else float(env)
        else:
            fd_const = getattr(ldm_stable, "pc_fd_const", None)
    return float(const if fd_const is None else fd_const)


def get_eigenvectors(ldm_stable, xt: torch.Tensor, text_emb: PromptEmbeddings, uncond_emb: PromptEmbeddings,
                     latents: torch.Tensor, mask: torch.Tensor, t: torch.Tensor, x0_pred: torch.Tensor,
                     pc_mode: PCStreamChoice = PCStreamChoice.BOTH, const: float = 1e-3, cfg_tar: float = 3,
                     iters: int = 50, double_precision: bool = False, eta: float = 1, n_ev: int = 1, group=None,
                     init_eigvecs: Optional[torch.Tensor] = None, fd_const: Optional[float] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor], List[torch.Tensor],
                                Dict[int, torch.Tensor], Dict[int, torch.Tensor]]:
    """Subspace (power) iteration on the Jacobian of the posterior mean (pc_drift.py:96-198).
    `group` (extension; None = the reference's single-process behaviour): process group over which the n_ev directions
    are sharded, see the module docstring.  `init_eigvecs` (extension): explicit start `[n_ev, C, H, W]` used instead of
    the `randn_like(xt) * mask * const` draw of :130 (seed-independent comparisons against the reference).
    `fd_const` (extension): the step of the finite difference actually taken.  The iteration only uses Ab / const
    (direction and eigenvalue), which is independent of the step to first order, but the reference default 1e-3 spreads a
    perturbation of 1e-3 / sqrt(D) per element — below the resolution of 16-bit tensor-core operands (and of the TF32
    matmuls the reference itself enables on a GPU, utils.py:116; its CPU fp32 run is the one that resolves it, DESIGN.md
    §2).  Resolution order: the argument; env AEDIT_PC_FD_CONST (a number, or "reference" for the caller's `const`);
    the evaluator's own `ldm_stable.pc_fd_const` (wrappers of models.py: 1.0 with fp16 operands, 8.0 with the bf16 build;
    absent = `const` for any other evaluator).  All outputs stay in units of the caller's `const`
    (tests/test_gpu_pc_drift.py::test_unet_jvp_resolves_with_fd_const)."""
    const_ret = const
    const = resolve_fd_const(ldm_stable, const, fd_const)
    from . import parallel as _par
    lib = _lib.load()
    dev = ldm_stable.device
    rank, ws = _par.world(group) if group is not None else (0, 1)
    rows = _par.shard_indices(n_ev, rank, ws) if ws > 1 else list(range(n_ev))
    xt1 = _f32(xt, dev)
    shape = (n_ev, *xt1.shape[1:])
    n_el = xt1[0].numel()
    x0_ref = _f32(x0_pred, dev)[0].contiguous()
    mask_d = _f32(mask, dev)
    mask1 = mask_d[0].expand(xt1.shape[1:]).contiguous()
    xt_n = xt1.expand(shape) if xt1.shape[0] == 1 else xt1
    # random start: randn_like of the n_ev-row tensor like the reference (:130), one draw for all ranks
    if init_eigvecs is not None:
        scaled = _f32(init_eigvecs, dev).reshape(shape).clone()
    else:
        scaled = torch.randn(shape, device=dev, dtype=torch.float32) * mask_d * const
    if ws > 1:
        _par.broadcast_(scaled, 0, group)
    prev = scaled.clone()
    cur = torch.empty_like(scaled)
    ws_bytes = int(lib.ae_pc_workspace_bytes(n_ev, n_el))
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    in_corr, in_norm = [], []
    interm_eigvecs, interm_eigvals = {}, {}
    sigma2 = ldm_stable.get_sigma(t) ** 2
    pick = (lambda v: v if (v is None or len(v) != n_ev) else v[rows])
    unc_r = PromptEmbeddings(*[pick(v) for v in uncond_emb])
    txt_r = PromptEmbeddings(*[pick(v) for v in text_emb])
    lat_r = pick(latents)
    with torch.no_grad():
        for i in range(iters):
            if ws == 1:
                _, x0p = forward_directional(ldm_stable, xt_n, t, latents, uncond_emb, text_emb, cfg_tar, eta=eta,
                                             eigvecs=scaled, amount=1, double_precision=double_precision, mode=pc_mode)
            else:
                local = None
                if rows:
                    _, local = forward_directional(ldm_stable, xt_n[rows], t, lat_r, unc_r, txt_r, cfg_tar, eta=eta,
                                                   eigvecs=scaled[rows], amount=1, double_precision=double_precision,
                                                   mode=pc_mode)
                x0p = _par.allgather_rows(local, n_ev, xt_n[0], group)
            norms = torch.empty(n_ev, device=dev, dtype=torch.float32)
            corr = torch.empty(n_ev, device=dev, dtype=torch.float32) if i > 0 else None
            new_scaled = torch.empty_like(scaled)
            _lib.check(lib.ae_pc_subspace_step(_p(x0p), _p(x0_ref), _p(mask1), _p(prev) if i > 0 else None, n_ev, n_el,
                                               float(const), _p(cur), _p(new_scaled), _p(norms), _p(corr), _p(work),
                                               ws_bytes, _stream()), "ae_pc_subspace_step")
            norm_of_Ab = norms if n_ev > 1 else norms[0]
            if i > 0:
                in_corr.append(corr)
            in_norm.append(norm_of_Ab)
            if not (i % 10) and i > 15:
                # the reference stores the tensor it then scales IN PLACE by `const` (:188-193): the stored
                # intermediate directions carry that factor
                interm_eigvecs[i] = new_scaled if const == const_ret else new_scaled * (const_ret / const)
                interm_eigvals[i] = norm_of_Ab / const * sigma2
            prev, cur = cur, prev
            scaled = new_scaled
    eigval = norm_of_Ab / const * sigma2                                               # :195
    eigvecs = scaled / const                                                            # :196 (`eigvecs /= const`)
    return eigvecs, eigval, in_corr, in_norm, interm_eigvecs, interm_eigvals


def apply_drift(ldm_stable, xt_m1: torch.Tensor, x0_pred: torch.Tensor, t: torch.Tensor, timesteps: torch.Tensor,
                num_diff_steps: int, eigdata: Dict[int, Dict[str, torch.Tensor]], latent: torch.Tensor,
                device: torch.device, use_shifted_x0_for_noisepred: bool = True,
                use_specific_ts_pc: Optional[int] = None, amount: float = 1, sub_iters: Optional[int] = None,
                eta: float = 1, ev_nums: List[int] = [1], evals: Optional[Dict[int, torch.Tensor]] = None
                ) -> torch.Tensor:
    """Shift the posterior mean along the stored principal directions and re-compose x_{t-1} (pc_drift.py:201-278)."""
    lib = _lib.load()
    # ---- which stored direction / eigenvalue (host look-ups, :219-231)
    use_t = t.item() if use_specific_ts_pc is None else timesteps[num_diff_steps - use_specific_ts_pc].item()
    if sub_iters is not None and evals is not None:
        raise ValueError("evals should be None if sub_iters is not None")
    if sub_iters is not None:
        eigvec = eigdata[use_t]['interm_eigvecs'][sub_iters].to(device)
        eigval = eigdata[t.item()]['interm_eigvals'][sub_iters].to(device)
    else:
        eigvec = eigdata[use_t]['eigvec'].to(device)
        eigval = eigdata[t.item()]['eigval'].to(device) if evals is None else torch.from_numpy(evals[t.item()]).to(device)
    shift_by = 0
    for ev_num in ev_nums:                                                                  # :232-235
        ev_idx = ev_num - 1
        shift_by = shift_by + amount * (eigval[ev_idx].unsqueeze(0).sqrt() * eigvec[ev_idx].unsqueeze(0))
    # ---- scalars of the step, evaluated on the host with the reference's expressions (:240-249)
    sched = ldm_stable.model.scheduler
    prev_timestep = t - sched.config.num_train_timesteps // sched.num_inference_steps
    variance = sched._get_variance(t, prev_timestep)
    std_dev_t = eta * variance ** (0.5)
    alpha_prod_t_prev = sched.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else sched.final_alpha_cumprod
    alpha_prod_t = sched.alphas_cumprod[int(t)]
    beta_prod_t = 1 - alpha_prod_t
    c_dir = (1 - alpha_prod_t_prev - std_dev_t ** 2) ** (0.5)
    ratio = (alpha_prod_t ** (0.5)) / (beta_prod_t ** (0.5))
    xm = _f32(xt_m1, device)
    x0p = _f32(x0_pred, device)
    if x0p.shape != xm.shape:
        x0p = x0p.expand_as(xm).contiguous()
    rows = xm.shape[0]
    n_el = xm[0].numel()
    sh = _f32(shift_by, device).reshape(-1)
    if sh.numel() != n_el:
        raise ValueError(f"eigvec has {sh.numel()} elements per direction, the latent {n_el}")
    lat = None
    if eta > 0:
        lat = _f32(latent, device)
        if lat.shape != xm.shape:
            lat = lat.expand_as(xm).contiguous()
    out = torch.empty_like(xm)
    _lib.check(lib.ae_pc_apply_drift(_p(xm), _p(x0p), _p(lat), _p(sh), float(std_dev_t), float(alpha_prod_t_prev ** (0.5)),
                                     float(c_dir), float(ratio), int(eta > 0), int(bool(use_shifted_x0_for_noisepred)),
                                     rows, n_el, _p(out), _stream()), "ae_pc_apply_drift")
    return out
