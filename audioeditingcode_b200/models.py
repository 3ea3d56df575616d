"""Drop-in replacement of the reference's model-wrapper boundary (code/models.py): same class names, method
names, argument meaning and error behaviour for the audio path — `PipelineWrapper`, `AudioLDMWrapper`,
`AudioLDM2Wrapper`, `TangoWrapper`, `load_model` — so the loops in ddm_inversion/inversion_utils.py and
pc_drift.py (and the reference's own main_run*.py) run unchanged on top of it.

What is different underneath: there is no diffusers pipeline object.  `self.model` is a light namespace exposing
exactly the attributes the callers touch (`.unet.config.in_channels`, `.scheduler`,
`.vocoder.config`, `.vae_scale_factor`), and every device computation goes to libaedit.so:
    unet_forward                       -> UNetEngine (tcgen05 GEMM/conv, fused norms, attention)      models.py:160-393
    sample_xts_from_x0                 -> ae_sample_xts                                                models.py:67-83
    get_zs_from_xts                    -> ae_cfg_inv_step (P = 0)                                      models.py:85-117
    reverse_step_with_custom_noise     -> ae_cfg_rev_step (P = 0)                                      models.py:119-158
Image / StableAudio wrappers of the reference (models.py:902-1354) are out of scope.
"""
from __future__ import annotations

import ctypes as C
import os
import types
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from . import _lib
from .scheduler import DDIMScheduler, SchedTable
from .unet import UNetEngine, TextCache
from .unet_config import UNetConfig, from_model_id
from . import weights as W


class UNet2DConditionOutput:
    """Stand-in for diffusers' output dataclass: the loops only read `.sample` (models.py:7,393)."""

    def __init__(self, sample: torch.Tensor):
        self.sample = sample


def _ver(t: torch.Tensor) -> int:
    """torch version counter of a tensor, or -1 for inference tensors (they carry none)."""
    try:
        return t._version
    except RuntimeError:
        return -1


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class PipelineWrapper(torch.nn.Module):
    family = "audioldm"
    # Finite-difference step pc_drift.get_eigenvectors takes through THIS evaluator when the caller does not pass
    # `fd_const`: the U-Net here runs on 16-bit tensor-core operands, which cannot resolve the
    # reference's const = 1e-3 (1e-3 / sqrt(D) per element); results are returned in units of the caller's `const`.
    # None = take the caller's `const` literally.  1.0 resolves through fp16 operands (11 significand bits; min cos 0.96 vs the
    # reference's fp32 finite differences through the tiny U-Net); the bf16 build (8 bits) needs 8x the step for the same
    # resolution (measured 0.59 at 1.0, 0.93 at 4.0).
    @property
    def pc_fd_const(self) -> Optional[float]:
        return getattr(self, "_pc_fd_const", 8.0 if _lib.operand_torch_dtype() == torch.bfloat16 else 1.0)

    @pc_fd_const.setter
    def pc_fd_const(self, value: Optional[float]) -> None:
        self._pc_fd_const = value

    def __init__(self, model_id: str, device: torch.device, double_precision: bool = False,
                 token: Optional[str] = None, *args, weights: Optional[Dict[str, torch.Tensor]] = None,
                 config: Optional[UNetConfig] = None, weight_seed: int = 0, allow_synthetic: Optional[bool] = None,
                 **kwargs) -> None:
        super().__init__()
        self.model_id = model_id
        # Seeded synthetic weights / text embeddings stand in for a checkpoint ONLY on explicit request (there is no
        # network here, so benchmarks and tests use them): `allow_synthetic=True`, AEDIT_ALLOW_SYNTHETIC=1, or a model id
        # under the reserved "synthetic/" namespace.  Otherwise a model id that is not a local checkpoint directory raises
        # (the reference would download it, models.py:478,556-564) instead of silently editing with random weights.
        if allow_synthetic is None:
            allow_synthetic = os.environ.get("AEDIT_ALLOW_SYNTHETIC", "0") != "0" or model_id.startswith("synthetic/")
        self.allow_synthetic = bool(allow_synthetic)
        self.device = torch.device(device)
        self.double_precision = double_precision
        self.token = token
        if double_precision:
            raise NotImplementedError("double_precision: the B200 path computes U-Net internals in bf16/fp32")
        cfg = config
        ckpt_dir = model_id if os.path.isdir(model_id) else None
        unet_path = None
        if ckpt_dir:
            # TANGO snapshots keep the U-Net inside pytorch_model_main.bin at the snapshot root (models.py:418-422) and
            # its architecture in the tango package; diffusers pipelines keep unet/config.json next to the weights
            root_main = os.path.join(ckpt_dir, "pytorch_model_main.bin")
            unet_path = root_main if (self.family == "tango" and os.path.exists(root_main)) else os.path.join(ckpt_dir, "unet")
        if cfg is None:
            cfg_json = os.path.join(ckpt_dir, "unet", "config.json") if ckpt_dir else None
            if cfg_json and os.path.exists(cfg_json):
                cfg = W.unet_config_from_json(cfg_json, os.path.basename(os.path.normpath(ckpt_dir)),
                                              scheduler_json=os.path.join(ckpt_dir, "scheduler", "scheduler_config.json"))
            elif ckpt_dir and self.family != "tango" and weights is None:
                # real weights must come with their architecture: a preset could silently disagree with the checkpoint
                # (e.g. the attention head split, which does not change any weight shape)
                raise FileNotFoundError(f"{cfg_json}: a checkpoint directory must carry unet/config.json")
            else:
                cfg = from_model_id(model_id)
        self.unet_config = cfg
        if weights is None:
            if ckpt_dir:
                weights = W.load_unet_checkpoint(unet_path, cfg)
                self.weights_source = f"checkpoint:{ckpt_dir}"
            elif self.allow_synthetic:
                weights = W.synthetic_weights(cfg, seed=weight_seed)
                self.weights_source = f"synthetic(seed={weight_seed})"
            else:
                raise FileNotFoundError(
                    f"{model_id!r} is not a local checkpoint directory (no network: hub ids cannot be downloaded). Pass a "
                    "directory laid out like the diffusers pipeline / TANGO snapshot the reference loads, or opt in to "
                    "seeded synthetic weights with allow_synthetic=True / AEDIT_ALLOW_SYNTHETIC=1.")
        else:
            self.weights_source = "caller"
        self.engine = UNetEngine(cfg, weights, self.device)
        unet_ns = types.SimpleNamespace(config=types.SimpleNamespace(in_channels=cfg.in_channels,
                                                                     sample_size=256, out_channels=cfg.out_channels),
                                        num_upsamplers=len(cfg.block_out_channels) - 1)
        vocoder_ns = types.SimpleNamespace(config=types.SimpleNamespace(model_in_dim=64, upsample_rates=[5, 4, 2, 2, 2],
                                                                        sampling_rate=16000))
        self.model = types.SimpleNamespace(unet=unet_ns, scheduler=None, vocoder=vocoder_ns, vae_scale_factor=4)
        self._text_cache: Dict[Any, TextCache] = {}
        self._sched_table: Optional[SchedTable] = None
        # a13 / f4: real tokenizer + text encoder(s) when the checkpoint directory carries them
        self.text_stack = None
        self._ckpt_dir = ckpt_dir
        from . import text_encoders as TE
        if TE.has_text_checkpoint(ckpt_dir, self.family):
            self.text_stack = TE.load_text_stack(ckpt_dir, self.family, self.device)

    # ---------------------------------------------------------------- scheduler plumbing
    @property
    def sched_table(self) -> SchedTable:
        return self.model.scheduler.table

    def get_sigma(self, timestep: int) -> float:                      # models.py:25-27
        sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.model.scheduler.alphas_cumprod - 1)
        return sqrt_recipm1_alphas_cumprod[timestep]

    def load_scheduler(self) -> None:
        cfg = self.unet_config
        self.model.scheduler = DDIMScheduler(cfg.beta_start, cfg.beta_end, prediction_type=cfg.prediction_type)

    def get_fn_STFT(self) -> torch.nn.Module:
        from .audio import TacotronSTFT
        return TacotronSTFT(filter_length=1024, hop_length=160, win_length=1024, n_mel_channels=64,
                            sampling_rate=16000, mel_fmin=0, mel_fmax=8000, device=self.device)

    def get_sr(self) -> int:
        return 16000

    def setup_extra_inputs(self, *args, **kwargs) -> None:           # models.py:47-48 (StableAudio only)
        pass

    def get_variance(self, timestep: torch.Tensor, prev_timestep: torch.Tensor) -> torch.Tensor:   # models.py:539-545
        alpha_prod_t = self.model.scheduler.alphas_cumprod[int(timestep)]
        alpha_prod_t_prev = self.get_alpha_prod_t_prev(prev_timestep)
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        return (beta_prod_t_prev / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)

    def get_alpha_prod_t_prev(self, prev_timestep: torch.Tensor) -> torch.Tensor:                 # models.py:547-549
        return self.model.scheduler.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 \
            else self.model.scheduler.final_alpha_cumprod

    def get_noise_shape(self, x0: torch.Tensor, num_steps: int) -> Tuple[int, ...]:                # models.py:60-65
        return (num_steps, self.model.unet.config.in_channels, x0.shape[-2], x0.shape[-1])

    # ---------------------------------------------------------------- a3: models.py:67-83
    def sample_xts_from_x0(self, x0: torch.Tensor, num_inference_steps: int = 50,
                           noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Samples from P(x_1:T|x_0).  `noise` ([N, C, H, W], draw order = ascending t like the reference's loop)
        may be passed explicitly; otherwise it is drawn with N torch.randn_like calls in the reference's order so
        a seeded run consumes the generator identically."""
        tab = self.sched_table
        N = tab.N
        assert N == num_inference_steps
        x0 = x0.to(self.device, torch.float32).contiguous()
        if noise is None:
            noise = torch.stack([torch.randn_like(x0[0]) for _ in range(N)])
        noise = noise.to(self.device, torch.float32).contiguous()
        xts = torch.empty(self.get_noise_shape(x0, N + 1), device=self.device, dtype=torch.float32)
        n_el = x0[0].numel()
        _lib.check(tab.lib.ae_sample_xts(tab.h, _ptr(x0), _ptr(noise), _ptr(xts), n_el, _stream()), "ae_sample_xts")
        return xts

    # ---------------------------------------------------------------- a4: models.py:85-117
    def get_zs_from_xts(self, xt: torch.Tensor, xtm1: torch.Tensor, noise_pred: torch.Tensor, t: torch.Tensor,
                        eta: float = 0, numerical_fix: bool = True, **kwargs
                        ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        tab = self.sched_table
        pos = tab.pos_of_t(int(t))
        tab.set_etas([eta] * tab.N)
        n_el = xt.numel()
        # the kernel addresses xt as xt_src[idx+1] and xtm1 as xts[idx]: hand it 2-row scratch views
        idx = tab.N - pos - 1
        xt = xt.to(torch.float32).contiguous()
        xtm1_out = xtm1.to(torch.float32).clone().contiguous()
        z = torch.empty_like(xtm1_out)
        base_src = xt.data_ptr() - (idx + 1) * n_el * 4
        base_xts = xtm1_out.data_ptr() - idx * n_el * 4
        base_zs = z.data_ptr() - idx * n_el * 4
        _lib.check(tab.lib.ae_cfg_inv_step(tab.h, pos, 1, float(eta), _ptr(noise_pred.to(torch.float32).contiguous()),
                                           n_el, None, 0, 0, None, C.c_void_p(base_src), C.c_void_p(base_xts),
                                           C.c_void_p(base_zs), int(bool(numerical_fix)), n_el, _stream()),
                   "ae_cfg_inv_step")
        return z, xtm1_out, None

    # ---------------------------------------------------------------- a5: models.py:119-158
    def reverse_step_with_custom_noise(self, model_output: torch.Tensor, timestep: torch.Tensor, sample: torch.Tensor,
                                       variance_noise: Optional[torch.Tensor] = None, eta: float = 0, **kwargs
                                       ) -> torch.Tensor:
        tab = self.sched_table
        pos = tab.pos_of_t(int(timestep))
        tab.set_etas([eta] * tab.N)
        if eta > 0 and variance_noise is None:
            variance_noise = torch.randn(model_output.shape, device=self.device)          # models.py:153-154
        out = torch.empty_like(sample, dtype=torch.float32)
        _lib.check(tab.lib.ae_cfg_rev_step(tab.h, pos, None, float(eta), _ptr(model_output.to(torch.float32).contiguous()),
                                           None, 0, None, _ptr(sample.to(torch.float32).contiguous()),
                                           _ptr(None if variance_noise is None else
                                                variance_noise.to(torch.float32).contiguous()),
                                           _ptr(out), None, None, None, sample.numel(), _stream()), "ae_cfg_rev_step")
        return out

    # ---------------------------------------------------------------- fused kernels (a4+a9, a5+a9)
    def k_cfg_inv_step(self, pos0: int, count: int, eta: float, eps_u, eps_c, P: int, cfg_map, xt_src, xts, zs,
                       numerical_fix: bool) -> None:
        """ae_cfg_inv_step over loop positions pos0..pos0+count-1 (see include/aedit.h).  eps_u: [count, ...],
        eps_c: [count*P, ...] (j-major), cfg_map: [P, ...]; xts / zs are updated in place."""
        tab = self.sched_table
        n_el = xts[0].numel()
        _lib.check(tab.lib.ae_cfg_inv_step(tab.h, pos0, count, float(eta), _ptr(eps_u), n_el,
                                           _ptr(eps_c) if P > 0 else None, n_el, P, _ptr(cfg_map) if P > 0 else None,
                                           _ptr(xt_src), _ptr(xts), _ptr(zs), int(bool(numerical_fix)), n_el,
                                           _stream()), "ae_cfg_inv_step")

    def k_cfg_rev_step(self, pos: int, eta: float, eps_u, eps_c, P: int, cfg_map, xt, z, out, masks=None,
                       fix_alpha=None, xT_fix=None, d_pos=None) -> None:
        """ae_cfg_rev_step at loop position pos (or *d_pos on device)."""
        tab = self.sched_table
        fix_h = None
        if fix_alpha is not None:
            fix_h = (C.c_float * 8)(*([float(v) for v in fix_alpha] + [0.0] * (8 - len(fix_alpha))))
        _lib.check(tab.lib.ae_cfg_rev_step(tab.h, pos, _ptr(d_pos), float(eta), _ptr(eps_u),
                                           _ptr(eps_c) if P > 0 else None, P, _ptr(cfg_map) if P > 0 else None,
                                           _ptr(xt), _ptr(z), _ptr(out), _ptr(masks) if fix_h is not None else None,
                                           C.cast(fix_h, C.c_void_p) if fix_h is not None else None,
                                           _ptr(xT_fix) if fix_h is not None else None, xt.numel(), _stream()),
                   "ae_cfg_rev_step")

    # ---------------------------------------------------------------- a7/a8: models.py:160-393, :691-899
    def _text_for(self, encoder_hidden_states, class_labels, encoder_attention_mask):
        """Maps the reference's (encoder_hidden_states, class_labels, encoder_attention_mask) triple onto the
        engine's text streams.  Overridden per family."""
        raise NotImplementedError

    def _cached_text(self, streams, masks) -> TextCache:
        key = tuple((None if s is None else (s.data_ptr(), tuple(s.shape), _ver(s))) for s in list(streams) + list(masks))
        tc = self._text_cache.get(key)
        if tc is None:
            if len(self._text_cache) > 16:
                self._text_cache.clear()
                self.engine.evict_graphs(self.live_texts())
            tc = self.engine.prepare_text(streams, masks)
            tc._keep = (streams, masks)
            self._text_cache[key] = tc
        return tc

    def live_texts(self):
        """TextCache objects still referenced by this wrapper's text caches (UNetEngine.evict_graphs keeps their graphs)."""
        live = list(self._text_cache.values())
        live += [h[0] for h in self.__dict__.get("_loop_text_cache", {}).values() if h[0] is not None]
        live += [h[0] for h in self.__dict__.get("_pair_text_cache", {}).values() if h[0] is not None]
        return live

    def unet_forward(self,
                     sample: torch.FloatTensor,
                     timestep: Union[torch.Tensor, float, int],
                     encoder_hidden_states: torch.Tensor,
                     class_labels: Optional[torch.Tensor] = None,
                     timestep_cond: Optional[torch.Tensor] = None,
                     attention_mask: Optional[torch.Tensor] = None,
                     cross_attention_kwargs: Optional[Dict[str, Any]] = None,
                     added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None,
                     down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                     mid_block_additional_residual: Optional[torch.Tensor] = None,
                     encoder_attention_mask: Optional[torch.Tensor] = None,
                     replace_h_space: Optional[torch.Tensor] = None,
                     replace_skip_conns: Optional[Dict[int, torch.Tensor]] = None,
                     return_dict: bool = True,
                     zero_out_resconns: Optional[Union[int, List]] = None) -> Tuple:
        if timestep_cond is not None or attention_mask is not None or down_block_additional_residuals is not None \
                or added_cond_kwargs is not None:
            raise NotImplementedError("timestep_cond / attention_mask / down_block_additional_residuals / "
                                      "added_cond_kwargs are never passed on the audio editing path")
        B = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.int64)
        t = timestep.reshape(-1).to(torch.int64)
        if t.numel() == 1:
            t = t.expand(B)
        streams, masks, cl = self._text_for(encoder_hidden_states, class_labels, encoder_attention_mask)
        text = slot = None
        if streams:
            text = self._cached_text(streams, masks)
            if text.n_rows == B:
                slot = torch.arange(B, dtype=torch.int32, device=self.device)
            elif text.n_rows == 1:
                slot = torch.zeros(B, dtype=torch.int32, device=self.device)
            else:
                raise ValueError(f"text batch {text.n_rows} does not match sample batch {B}")
        if cl is not None and cl.shape[0] != B:
            cl = cl.expand(B, -1)
        out, h_space, extracted = self.engine.forward(
            sample, t, text=text, slot_map=slot, class_labels=cl,
            mid_block_additional_residual=mid_block_additional_residual, replace_h_space=replace_h_space,
            replace_skip_conns=replace_skip_conns, zero_out_resconns=zero_out_resconns, want_taps=True)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out), h_space, extracted

    def _pair_setup(self, n: int, uncond, cond):
        """Text rows / slot map of a CFG pair batch [n uncond rows | n cond rows], cached per embedding tensors."""
        from .ddm_inversion import inversion_utils as IU
        flat = [v for tr in (uncond, cond) for v in tr]
        key = ("pair", n) + tuple(None if v is None else (v.data_ptr(), tuple(v.shape), _ver(v)) for v in flat)
        cache = self.__dict__.setdefault("_pair_text_cache", {})
        hit = cache.get(key)
        if hit is None:
            if len(cache) > 8:
                cache.clear()
                self.engine.evict_graphs(self.live_texts())
            streams, masks, cl = IU._cat_text(self, uncond, cond)
            ru = next((v.shape[0] for v in uncond if v is not None), 1)
            rc = next((v.shape[0] for v in cond if v is not None), 1)
            if ru not in (1, n) or rc not in (1, n):
                raise ValueError(f"text batches {ru} / {rc} do not match the sample batch {n}")
            slot = torch.cat([torch.arange(n, dtype=torch.int32) if ru == n else torch.zeros(n, dtype=torch.int32),
                              ru + (torch.arange(n, dtype=torch.int32) if rc == n else torch.zeros(n, dtype=torch.int32))]
                             ).to(self.device)
            text = self.engine.prepare_text(streams, masks) if streams else None
            cl_rows = None if cl is None else cl.index_select(0, slot.long())
            hit = (text, cl_rows, slot, (ru, rc), (flat, streams, masks))
            cache[key] = hit
        return hit

    def cfg_pair_eval_batch(self, x: torch.Tensor, timestep, uncond, cond) -> torch.Tensor:
        """eps of a prebuilt CFG pair batch x = [n rows evaluated under `uncond` | n rows under `cond`] as ONE batched,
        CUDA-graph-cached U-Net evaluation (pc_drift.py:64-80 issues two eager calls).  uncond / cond: (encoder_hidden_states,
        class_labels, encoder_attention_mask) triples as returned by encode_text, with 1 or n rows each."""
        from .ddm_inversion import inversion_utils as IU
        n = x.shape[0] // 2
        text, cl_rows, slot, (ru, rc), _ = self._pair_setup(n, uncond, cond)
        if not torch.is_tensor(timestep):
            timestep = torch.tensor(timestep)
        t_in = timestep.reshape(-1)[:1].to(self.device, torch.int64).expand(2 * n).contiguous()
        return IU._unet_eval(self, x.to(self.device, torch.float32), t_in, text, slot, cl_rows, slot_key=("pair", n, ru, rc))

    def cfg_pair_eval(self, x_u: torch.Tensor, x_c: torch.Tensor, timestep, uncond, cond
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """eps(x_u | uncond) and eps(x_c | cond) — what two `unet_forward` calls return as `.sample` (ddim_inversion.py:23-41)."""
        n = x_u.shape[0]
        eps = self.cfg_pair_eval_batch(torch.cat([x_u, x_c], 0), timestep, uncond, cond)
        return eps[:n], eps[n:]

    # ---------------------------------------------------------------- text (a13)
    def _encode_text_or_synthetic(self, prompts: List[str], synthetic):
        """The checkpoint's tokenizer / text encoder(s) (text_encoders.py); seeded synthetic embeddings only when the
        wrapper was explicitly created with allow_synthetic (benchmarks / tests without a checkpoint)."""
        if self.text_stack is not None:
            return self.text_stack(list(prompts))
        if not self.allow_synthetic:
            raise FileNotFoundError(
                f"{self.model_id!r}: no tokenizer / text encoder found in the checkpoint directory and synthetic text "
                "embeddings were not requested (allow_synthetic=True / AEDIT_ALLOW_SYNTHETIC=1)")
        return synthetic(list(prompts))

    def _synthetic_text(self, prompts: List[str], dim: int, L: Optional[int], normalize: bool, salt: int):
        """Deterministic stand-in embeddings when no text-encoder checkpoint is on disk (no network here):
        N(0,1) seeded by the prompt string."""
        rows = []
        for p in prompts:
            seed = (sum((i + 1) * ord(c) for i, c in enumerate(p)) + 7919 * salt) % (2 ** 31)
            g = torch.Generator().manual_seed(1000 + seed)
            n = 1 if L is None else L
            e = torch.randn(n, dim, generator=g)
            if normalize:
                e = torch.nn.functional.normalize(e, dim=-1)
            rows.append(e)
        return torch.stack(rows).to(self.device)

    # ---------------------------------------------------------------- ends (a10-a12), filled by ends.py
    def _ends(self):
        if getattr(self, "_ends_obj", None) is None:
            from .ends import AudioEnds
            self._ends_obj = AudioEnds(self.device, self._ckpt_dir, allow_synthetic=self.allow_synthetic)
        return self._ends_obj

    def vae_encode(self, x: torch.Tensor) -> torch.Tensor:                                    # models.py:495-499
        if x.shape[2] % 4:
            x = torch.nn.functional.pad(x, (0, 0, 4 - (x.shape[2] % 4), 0))
        return self._ends().vae_encode_mode(x).float()

    def vae_decode(self, x: torch.Tensor) -> torch.Tensor:                                    # models.py:502-503
        return self._ends().vae_decode(x)

    def decode_to_mel(self, x: torch.Tensor) -> torch.Tensor:                                 # models.py:505-509
        return self._ends().vocoder(x[0, 0].detach().float()).detach().unsqueeze(0)


class AudioLDMWrapper(PipelineWrapper):
    """models.py:475-549.  Conditioning = L2-normalised 512-d CLAP text embedding passed as `class_labels`."""

    def _text_for(self, encoder_hidden_states, class_labels, encoder_attention_mask):
        return [], [], class_labels

    def encode_text(self, prompts: List[str], **kwargs) -> Tuple[None, Optional[torch.Tensor], None]:   # :511-537
        return self._encode_text_or_synthetic(
            prompts, lambda ps: (None, self._synthetic_text(ps, 512, None, True, 0)[:, 0], None))


class AudioLDM2Wrapper(PipelineWrapper):
    """models.py:552-899.  encode_text returns (GPT-2 generated [P,8,768], T5 [P,L,1024], T5 mask [P,L]);
    unet_forward routes them as stream 0 (unmasked) and stream 1 (masked) — the translation of models.py:706-710."""

    def _text_for(self, encoder_hidden_states, class_labels, encoder_attention_mask):
        return [encoder_hidden_states, class_labels], [None, encoder_attention_mask], None

    family = "audioldm2"

    def encode_text(self, prompts: List[str], **kwargs) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:   # :599-677
        return self._encode_text_or_synthetic(prompts, self._synthetic_triple)

    def _synthetic_triple(self, prompts: List[str]):
        L = max(1, max(len(p.split()) for p in prompts) + (1 if any(p for p in prompts) else 0))
        dims = {sp[1]: sp[0] for sp in self.unet_config.transformer_specs if sp is not None}     # 768 / 1024 at full size
        gen = self._synthetic_text(prompts, dims.get(0, 768), 8, False, 1)
        t5 = self._synthetic_text(prompts, dims.get(1, 1024), L, False, 2)
        mask = torch.zeros(len(prompts), L, dtype=torch.long, device=self.device)
        for i, p in enumerate(prompts):
            mask[i, : max(1, len(p.split()) + (1 if p else 0))] = 1
        return gen, t5, mask

    def decode_to_mel(self, x: torch.Tensor) -> torch.Tensor:                                 # models.py:591-597
        tmp = self._ends().vocoder(x[:, 0].detach().float()).detach()
        if len(tmp.shape) == 1:
            tmp = tmp.unsqueeze(0)
        return tmp


class TangoWrapper(PipelineWrapper):
    """models.py:396-472.  Conditioning = T5 hidden states + boolean mask; v-prediction scheduler (SD-2.1 betas)."""

    def _text_for(self, encoder_hidden_states, class_labels, encoder_attention_mask):
        return [encoder_hidden_states], [encoder_attention_mask], None

    def vae_encode(self, x: torch.Tensor) -> torch.Tensor:                                    # models.py:439-447
        if x.shape[2] % 4:
            x = torch.nn.functional.pad(x, (0, 0, 4 - (x.shape[2] % 4), 0))
        if x.shape[2] > 1700:
            raise RuntimeWarning("This model dies at this point")
        return self._ends().vae_encode_sample(x).float()

    family = "tango"

    def decode_to_mel(self, x: torch.Tensor):                                                 # models.py:452-453
        """tango's AutoencoderKL.decode_to_waveform (in-tree twin autoencoder.py:63-66 -> hifigan/utilities.py:76-85):
        vocoder on [B, T, 64], then `(wav * 32768).astype(int16)` — returned as an int16 numpy array [B, samples]; the
        conversion runs on the device (ae_wave_to_int16), only the PCM samples cross to the host."""
        wav = self._ends().vocoder(x[:, 0].detach().float())
        if wav.dim() == 1:
            wav = wav.unsqueeze(0)
        wav = wav.contiguous()
        pcm = torch.empty(wav.shape, dtype=torch.int16, device=wav.device)
        self.engine.ops.wave_to_int16(wav, pcm)
        return pcm.cpu().numpy()

    def encode_text(self, prompts: List[str], **kwargs) -> Tuple[Optional[torch.Tensor], None, Optional[torch.Tensor]]:
        return self._encode_text_or_synthetic(prompts, self._synthetic_triple)                               # :455-460

    def _synthetic_triple(self, prompts: List[str]):
        L = max(1, max(len(p.split()) for p in prompts) + 1)
        dims = {sp[1]: sp[0] for sp in self.unet_config.transformer_specs if sp is not None}
        t5 = self._synthetic_text(prompts, dims.get(0, 1024), L, False, 3)
        mask = torch.zeros(len(prompts), L, dtype=torch.bool, device=self.device)
        for i, p in enumerate(prompts):
            mask[i, : len(p.split()) + 1] = True
        return t5, None, mask


def load_model(model_id: str, device: torch.device, num_diffusion_steps: int,
               double_precision: bool = False, token: Optional[str] = None, **kwargs) -> PipelineWrapper:
    """models.py:1357-1374: substring dispatch, load_scheduler(), scheduler.set_timesteps(N)."""
    if 'tango' in model_id:
        ldm_stable = TangoWrapper(model_id=model_id, device=device, double_precision=double_precision, token=token, **kwargs)
    elif 'audioldm2' in model_id:
        ldm_stable = AudioLDM2Wrapper(model_id=model_id, device=device, double_precision=double_precision, token=token, **kwargs)
    elif 'audioldm' in model_id:
        ldm_stable = AudioLDMWrapper(model_id=model_id, device=device, double_precision=double_precision, token=token, **kwargs)
    else:
        raise ValueError(f"{model_id}: only the AudioLDM / AudioLDM2 / TANGO audio path is implemented "
                         "(image and Stable Audio wrappers of the reference are out of scope)")
    ldm_stable.load_scheduler()
    ldm_stable.model.scheduler.set_timesteps(num_diffusion_steps, device=device)
    return ldm_stable
