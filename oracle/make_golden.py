"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py).  Run in the build container:

    python -m oracle.make_golden

What runs reference code here:
  * code/ddm_inversion/inversion_utils.py  inversion_forward_process / inversion_reverse_process   (unmodified)
  * code/models.py                         PipelineWrapper.{sample_xts_from_x0,get_zs_from_xts,
                                           reverse_step_with_custom_noise}, AudioLDMWrapper.{get_variance,
                                           get_alpha_prod_t_prev}                                   (unmodified)
  * code/audioldm/latent_diffusion/openaimodel.py  UNetModel (the only in-tree U-Net)               (unmodified)
  * code/audioldm/audio/stft.py            TacotronSTFT.mel_spectrogram                             (unmodified)
  * code/audioldm/variational_autoencoder/modules.py Encoder / Decoder, code/audioldm/hifigan/models.py Generator
What is substituted (absent third-party code): the diffusers scheduler object -> oracle.ddpm_oracle.MiniDDIM,
the diffusers U-Net module tree -> the vendored UNetModel called from an overridden `unet_forward`,
text encoders -> fixed seeded embedding vectors.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import, unet_torch as U          # noqa: E402
from oracle.ddpm_oracle import MiniDDIM                  # noqa: E402
from audioeditingcode_b200 import unet_config as C      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def prompt_vector(prompt: str, dim: int = 512) -> torch.Tensor:
    """Deterministic stand-in for the CLAP text projection: L2-normalised N(0,1) seeded by the prompt."""
    seed = sum((i + 1) * ord(c) for i, c in enumerate(prompt)) % (2 ** 31)
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.nn.functional.normalize(torch.randn(1, dim, generator=g), dim=-1)


def build_ldm_unet(ref, cfg, weights):
    assert cfg.transformer_specs == (None,)
    ch = cfg.block_out_channels
    mult = [c // ch[0] for c in ch]
    att = [2 ** i for i, a in enumerate(cfg.attn_levels) if a]
    m = ref.openaimodel.UNetModel(
        image_size=64, extra_film_condition_dim=cfg.class_embed_dim, extra_film_use_concat=True,
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, model_channels=ch[0],
        attention_resolutions=att, num_res_blocks=cfg.layers_per_block, channel_mult=mult,
        num_head_channels=ch[-1] // cfg.num_heads[-1], use_spatial_transformer=True).eval()
    m.load_state_dict(U.canonical_to_ldm(cfg, weights))
    return m


def make_fake_wrapper(ref, cfg, weights, n_steps, prediction_type="epsilon"):
    """Reference AudioLDMWrapper with __init__ bypassed: the loop/scheduler math is the
    reference's own; only the absent diffusers objects are substituted."""
    unet = build_ldm_unet(ref, cfg, weights)

    class Fake(ref.models.AudioLDMWrapper):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.model_id = "fake/audioldm"
            self.device = torch.device("cpu")
            self.double_precision = False
            sched = MiniDDIM(cfg.beta_start, cfg.beta_end, prediction_type=prediction_type)
            sched.set_timesteps(n_steps)
            self.model = types.SimpleNamespace(
                unet=types.SimpleNamespace(config=types.SimpleNamespace(in_channels=cfg.in_channels)),
                scheduler=sched)
            self.calls = 0
            self.rec_fwd, self.rec_rev, self.rec_unet = [], [], []

        # record the noise predictions the reference hands to its scheduler math, then run that math unmodified
        def get_zs_from_xts(self, xt, xtm1, noise_pred, t, **kw):
            self.rec_fwd.append(noise_pred.clone())
            return super().get_zs_from_xts(xt, xtm1, noise_pred, t, **kw)

        def reverse_step_with_custom_noise(self, model_output, timestep, sample, **kw):
            self.rec_rev.append(model_output.clone())
            return super().reverse_step_with_custom_noise(model_output, timestep, sample, **kw)

        def encode_text(self, prompts, **kw):
            return None, torch.cat([prompt_vector(p) for p in prompts], 0), None

        def unet_forward(self, sample, timestep, encoder_hidden_states=None, class_labels=None, **kw):
            self.calls += 1
            t = timestep if torch.is_tensor(timestep) else torch.tensor(timestep)
            t = t.reshape(-1).expand(sample.shape[0])
            with torch.no_grad():
                out = unet(sample, t, y=class_labels)
            self.rec_unet.append(out.clone())
            return ref_import.UNet2DConditionOutput(sample=out), None, None

    return Fake()


def run_loops(ref, cfg, weights, n_steps, H, W, src, tgt, cfg_src, cfg_tar, tstart, pred, seed, cutoff=None):
    model = make_fake_wrapper(ref, cfg, weights, n_steps, pred)
    g = torch.Generator().manual_seed(seed)
    x0 = 0.5 * torch.randn(1, cfg.in_channels, H, W, generator=g)
    # replicate the reference's N randn_like draws (models.py:79-81) to hand them to the port explicitly
    torch.manual_seed(seed + 1)
    noise = torch.stack([torch.randn_like(x0)[0] for _ in range(n_steps)])
    torch.manual_seed(seed + 1)
    with torch.no_grad():
        xt, zs, xts, _ = ref.inversion_utils.inversion_forward_process(
            model, x0, etas=1.0, prompts=list(src), cfg_scales=list(cfg_src), num_inference_steps=n_steps,
            numerical_fix=True, cutoff_points=cutoff)
        unet_fwd = model.rec_unet
        model.rec_unet = []
        ts = torch.tensor(tstart, dtype=torch.int)
        skip = n_steps - ts
        w_edit, _ = ref.inversion_utils.inversion_reverse_process(
            model, xT=xts, tstart=ts, etas=1.0, prompts=list(tgt), neg_prompts=[""], cfg_scales=list(cfg_tar),
            zs=zs[:int(n_steps - min(skip))], cutoff_points=cutoff)
    bf16_zs, bf16_edit = bf16_autocast_errors(cfg, weights, n_steps, pred, x0, noise, zs, xts, w_edit, src, tgt,
                                              cfg_src, cfg_tar, tstart, cutoff)
    return dict(x0=x0, noise=noise, zs=zs, xts=xts, w_edit=w_edit, bf16_autocast_err_zs=bf16_zs,
                bf16_autocast_err_edit=bf16_edit,
                eps_fwd=torch.cat(model.rec_fwd), eps_rev=torch.cat(model.rec_rev),
                # raw U-Net outputs of every step in loop order: [uncond (1 row), cond (P rows)] alternating
                eps_u_fwd=torch.cat(unet_fwd[0::2]), eps_c_fwd=torch.stack(unet_fwd[1::2]),
                eps_u_rev=torch.cat(model.rec_unet[0::2]), eps_c_rev=torch.stack(model.rec_unet[1::2]),
                uncond=prompt_vector(""), src=torch.cat([prompt_vector(p) for p in src]),
                tgt=torch.cat([prompt_vector(p) for p in tgt]))


def bf16_autocast_errors(cfg, weights, n_steps, pred, x0, noise, zs_ref, xts_ref, w_ref, src, tgt, cfg_src, cfg_tar,
                         tstart, cutoff):
    """Error level of STOCK PyTorch bf16 autocast on the same loops (oracle loop code, U-Net under
    torch.autocast(bfloat16)) relative to the fp32 reference: the yardstick for the bf16 tensor-core path's
    end-to-end tolerance (a bf16 path cannot be expected to beat it by much; ours keeps an fp32 residual stream)."""
    from oracle import ddpm_oracle as D
    sched = MiniDDIM(cfg.beta_start, cfg.beta_end, prediction_type=pred)
    sched.set_timesteps(n_steps)
    un = prompt_vector("")

    def mk(prompts):
        y_c = torch.cat([prompt_vector(p) for p in prompts])

        def fn(x, t, which):
            y = un if which == "uncond" else y_c
            tt = torch.full((x.shape[0],), int(t), dtype=torch.int64)
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                return U.unet_forward(cfg, weights, x, tt, class_labels=y)[0].float()
        return fn
    P = len(src)
    _, zs, xts = D.inversion_forward_process(sched, mk(src), x0, noise, 1.0, P, list(cfg_src), cutoff_points=cutoff,
                                             prompts=list(src))
    ts = torch.tensor(tstart, dtype=torch.int)
    w = D.inversion_reverse_process(sched, mk(tgt), xts, zs[:int(ts.max())], ts, 1.0, P, list(cfg_tar),
                                    cutoff_points=cutoff)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    return rel(zs, zs_ref), rel(w, w_ref)


def save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print("wrote", name, {k: tuple(v.shape) for k, v in out.items()})


def main():
    ref = ref_import.load()
    torch.set_num_threads(8)
    cfg = C.preset("tiny-audioldm")
    w = U.synthetic_weights(cfg, seed=0)
    H, W, N = 16, 16, 10

    # (1) single U-Net evaluation by the vendored UNetModel
    unet = build_ldm_unet(ref, cfg, w)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 8, H, W, generator=g)
    t = torch.tensor([981, 501, 1])
    y = torch.cat([prompt_vector(p) for p in ("", "a dog barking", "piano")])
    with torch.no_grad():
        eps = unet(x, t, y=y)
    save("unet_tiny_audioldm.npz", x=x, t=t, y=y, eps=eps, weight_seed=0)

    # (2) loops, eps-prediction, single prompt
    r = run_loops(ref, cfg, w, N, H, W, ["a recording of a dog"], ["a recording of a cat"], [3.0], [12.0],
                  [N // 2 + 2], "epsilon", seed=21)
    save("loop_eps_single.npz", n_steps=N, tstart=[N // 2 + 2], cfg_src=[3.0], cfg_tar=[12.0], **r)

    # (3) loops, v-prediction (TANGO scheduler branch of models.py:92-93,104-105,133-134,144-145)
    r = run_loops(ref, cfg, w, N, H, W, ["rain"], ["thunder"], [1.0], [3.0], [N], "v_prediction", seed=31)
    save("loop_vpred_single.npz", n_steps=N, tstart=[N], cfg_src=[1.0], cfg_tar=[3.0], **r)

    # (4) multi-prompt with per-prompt tstart (mask-fix branch inversion_utils.py:308-315) — H=32 so the
    #     15-tap blur's reflect padding (needs H, W > 7) is legal
    r = run_loops(ref, cfg, w, N, 32, W, ["speech", "music"], ["a choir", "a violin"], [3.0, 2.0], [5.0, 8.0],
                  [8, 6], "epsilon", seed=41)
    save("loop_eps_multi.npz", n_steps=N, tstart=[8, 6], cfg_src=[3.0, 2.0], cfg_tar=[5.0, 8.0], **r)

    # (5) uncond-only forward (prompts == [""], inversion_utils.py:86,110-111)
    model = make_fake_wrapper(ref, cfg, w, N)
    g = torch.Generator().manual_seed(51)
    x0 = 0.5 * torch.randn(1, 8, H, W, generator=g)
    torch.manual_seed(52)
    noise = torch.stack([torch.randn_like(x0)[0] for _ in range(N)])
    torch.manual_seed(52)
    with torch.no_grad():
        _, zs, xts, _ = ref.inversion_utils.inversion_forward_process(
            model, x0, etas=1.0, prompts=[""], cfg_scales=[3.5], num_inference_steps=N, numerical_fix=True)
    assert model.calls == N
    save("loop_eps_uncond_only.npz", n_steps=N, x0=x0, noise=noise, zs=zs, xts=xts, uncond=prompt_vector(""))

    # (6) scheduler scalars through the reference's own get_variance / get_alpha_prod_t_prev
    for n in (50, 100, 200):
        model = make_fake_wrapper(ref, cfg, w, n)
        ts = model.model.scheduler.timesteps
        var, ap = [], []
        for tt in ts:
            prev = tt - 1000 // n
            var.append(float(model.get_variance(tt, prev)))
            ap.append(float(model.get_alpha_prod_t_prev(prev)))
        save(f"sched_{n}.npz", timesteps=ts, variance=np.asarray(var, np.float32),
             alpha_prod_t_prev=np.asarray(ap, np.float32),
             alphas_cumprod=model.model.scheduler.alphas_cumprod)


def synthetic_wave(n, seed=3):
    """5 sines + noise, peak-normalised to 0.5 like tools.py:46-64 (SURVEY.md §8d synthetic audio)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float64) / 16000.0
    freqs = [110.0, 440.0, 1250.0, 3100.0, 6400.0]
    w = sum((0.6 ** i) * torch.sin(2 * np.pi * f * t + i) for i, f in enumerate(freqs))
    w = w + 0.05 * torch.randn(n, generator=g, dtype=torch.float64)
    w = w - w.mean()
    w = 0.5 * w / w.abs().max()
    return w.float()


def golden_stft():
    """Reference TacotronSTFT (audioldm/audio/stft.py, unmodified; librosa.filters.mel stubbed by the slaney
    restatement in oracle/ref_import.py) on a synthetic clip."""
    ref = ref_import.load()
    fn = ref.stft.TacotronSTFT(1024, 160, 1024, 64, 16000, 0, 8000)
    wav = synthetic_wave(16000 * 2 + 37)
    with torch.no_grad():
        mel, logmag, energy = fn.mel_spectrogram(wav[None])
    save("stft_mel.npz", wav=wav, mel=mel[0], logmag=logmag[0], energy=energy[0], mel_basis=fn.mel_basis,
         window=torch.from_numpy(np.asarray(__import__("scipy.signal").signal.get_window("hann", 1024, fftbins=True))).float())


def golden_ends():
    """Vendored VAE Encoder/Decoder (variational_autoencoder/modules.py) and HiFi-GAN Generator (hifigan/models.py),
    unmodified, on seeded synthetic weights (oracle/ends_torch.py naming converted by vae_to_ldm / hifigan_to_ldm)."""
    from oracle import ends_torch as E
    ref = ref_import.load()
    dd = dict(double_z=True, z_channels=8, resolution=256, downsample_time=False, in_channels=1, out_ch=1, ch=128,
              ch_mult=[1, 2, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    enc, dec = ref.vae_modules.Encoder(**dd).eval(), ref.vae_modules.Decoder(**dd).eval()
    w = E.vae_synthetic_weights(0)
    ld = E.vae_to_ldm(w)
    enc.load_state_dict({k[8:]: v for k, v in ld.items() if k.startswith("encoder.")})
    dec.load_state_dict({k[8:]: v for k, v in ld.items() if k.startswith("decoder.")})
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 1, 128, 64, generator=g) * 2 - 4          # log-mel-like
    with torch.no_grad():
        mom = torch.nn.functional.conv2d(enc(x), ld["quant_conv.weight"], ld["quant_conv.bias"])
        z = mom[:, :8] * E.VAE_SCALING
        dx = dec(torch.nn.functional.conv2d(z / E.VAE_SCALING, ld["post_quant_conv.weight"], ld["post_quant_conv.bias"]))
    rel = lambda a, b: float((a - b).norm() / b.norm())
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):     # stock torch bf16: the error yardstick
        e_enc = rel(E.vae_encode_moments(w, x).float(), mom)
        e_dec = rel(E.vae_decode(w, z).float(), dx)
    save("vae_ends.npz", x=x, moments=mom, z=z, decoded=dx, weight_seed=0, bf16_autocast_err_encode=e_enc,
         bf16_autocast_err_decode=e_dec)
    h = types.SimpleNamespace(resblock="1", upsample_rates=[5, 4, 2, 2, 2], upsample_kernel_sizes=[16, 16, 8, 4, 4],
                              upsample_initial_channel=1024, resblock_kernel_sizes=[3, 7, 11],
                              resblock_dilation_sizes=[[1, 3, 5]] * 3, num_mels=64)
    gen = ref.hifigan.Generator(h).eval()
    gen.remove_weight_norm()
    hw = E.hifigan_synthetic_weights(0)
    gen.load_state_dict(E.hifigan_to_ldm(hw))
    mel = torch.randn(1, 32, 64, generator=g) * 2 - 4
    with torch.no_grad():
        wav = gen(mel.transpose(1, 2)).squeeze(1)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        e_voc = rel(E.hifigan_forward(hw, mel).float(), wav)
    save("hifigan_ends.npz", mel=mel, wav=wav, weight_seed=0, bf16_autocast_err=e_voc)


def golden_ddim():
    """Reference DDIM baseline (code/ddm_inversion/ddim_inversion.py, unmodified): ddim_inversion + text2image_ldm_stable
    on the fake wrapper; the diffusers scheduler.step is substituted by oracle MiniDDIM.step ([UPSTREAM] restatement)."""
    ref = ref_import.load()
    cfg = C.preset("tiny-audioldm")
    w = U.synthetic_weights(cfg, seed=0)
    N = 10
    model = make_fake_wrapper(ref, cfg, w, N)
    # ddim_inversion.py:23-41 calls model.unet_forward with the encode_text() tuples
    g = torch.Generator().manual_seed(61)
    w0 = 0.5 * torch.randn(1, 8, 16, 16, generator=g)
    import ddm_inversion.ddim_inversion as DI
    with torch.no_grad():
        wT = DI.ddim_inversion(model, w0, ["a dog"], 3.0, num_inference_steps=N, skip=0)
        wrec = DI.text2image_ldm_stable(model, ["a cat"], N, 5.0, wT, skip=0)
    # stock torch bf16-autocast yardstick on the same two loops
    orig_fwd = model.unet_forward

    def bf16_fwd(*a, **k):
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out, h, e = orig_fwd(*a, **k)
        out.sample = out.sample.float()
        return out, h, e
    model.unet_forward = bf16_fwd
    with torch.no_grad():
        wT_b = DI.ddim_inversion(model, w0, ["a dog"], 3.0, num_inference_steps=N, skip=0)
        wrec_b = DI.text2image_ldm_stable(model, ["a cat"], N, 5.0, wT, skip=0)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    save("ddim_mode.npz", w0=w0, wT=wT, w_rec=wrec, n_steps=N, uncond=prompt_vector(""), src=prompt_vector("a dog"),
         tgt=prompt_vector("a cat"), bf16_autocast_err_inv=rel(wT_b, wT), bf16_autocast_err_rec=rel(wrec_b, wrec))


# ------------------------------------------------------------------------------------------------------------------
# pc_drift (code/pc_drift.py, unmodified) on a LINEAR denoiser with a known Jacobian spectrum
# ------------------------------------------------------------------------------------------------------------------
def linear_denoiser_factors(D, seed, lam_u=(0.9, 0.65, 0.45, 0.3, 0.2, 0.1), tilt=0.05):
    """Low-rank symmetric maps S_u = U diag(lam_u) U^T (unconditional) and S_c = U diag(lam_c) U^T (conditional):
    the posterior mean of the fake model is x0_hat = S x / sqrt(alpha_bar_t), i.e. eps = (x - S x) / sqrt(1 - alpha_bar_t)."""
    g = torch.Generator().manual_seed(seed)
    r = len(lam_u)
    U_, _ = torch.linalg.qr(torch.randn(D, r, generator=g, dtype=torch.float64))
    lam_u = torch.tensor(lam_u, dtype=torch.float64)
    lam_c = lam_u * (1 + tilt * torch.linspace(-1, 1, r, dtype=torch.float64))
    return U_.float(), lam_u.float(), lam_c.float()


class LinearLDM:
    """Duck-typed PipelineWrapper for the reference's pc_drift functions: linear denoiser + MiniDDIM scheduler."""

    def __init__(self, shape, n_steps, factors):
        self.device = torch.device("cpu")
        self.shape = shape
        self.U, self.lam_u, self.lam_c = factors
        sch = MiniDDIM(0.0015, 0.0195)
        sch.set_timesteps(n_steps)
        self.model = types.SimpleNamespace(scheduler=sch)

    def get_sigma(self, timestep):                      # models.py:25-27
        return torch.sqrt(1.0 / self.model.scheduler.alphas_cumprod - 1)[timestep]

    def eps(self, x, t, cond: bool):
        ab = self.model.scheduler.alphas_cumprod[int(t)]
        xf = x.reshape(x.shape[0], -1)
        lam = self.lam_c if cond else self.lam_u
        sx = ((xf @ self.U) * lam) @ self.U.T
        return ((xf - sx) / (1 - ab) ** 0.5).reshape(x.shape)

    def unet_forward(self, sample, timestep, encoder_hidden_states=None, class_labels=None, encoder_attention_mask=None,
                     **kw):
        cond = bool(class_labels is not None and float(class_labels.flatten()[0]) > 0.5)
        return types.SimpleNamespace(sample=self.eps(sample, timestep, cond)), None, None


def golden_pc_drift():
    ref = ref_import.load()
    PC = ref.pc_drift
    C_, H_, W_ = 8, 4, 16
    D = C_ * H_ * W_
    N = 20
    cases = [dict(name="a", n_ev=3, iters=12, seed=0, patch=None, mode="BOTH"),
             dict(name="b", n_ev=4, iters=12, seed=1, patch=(1, 3), mode="TEXT"),
             dict(name="c", n_ev=1, iters=8, seed=2, patch=None, mode="BOTH"),
             dict(name="d", n_ev=8, iters=21, seed=3, patch=None, mode="BOTH",
                  lam=(0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.32, 0.25, 0.18, 0.1))]     # rank >= n_ev: no null directions
    out = {}
    for cs in cases:
        fac = linear_denoiser_factors(D, 100 + cs["seed"], **({"lam_u": cs["lam"]} if "lam" in cs else {}))
        ldm = LinearLDM((C_, H_, W_), N, fac)
        g = torch.Generator().manual_seed(cs["seed"])
        xt = torch.randn(1, C_, H_, W_, generator=g)
        lat = torch.randn(1, C_, H_, W_, generator=g)
        mask = torch.zeros(1, C_, H_, W_)
        if cs["patch"] is None:
            mask[:] = 1
        else:
            mask[:, :, cs["patch"][0]:cs["patch"][1], :] = 1
        t = ldm.model.scheduler.timesteps[7]
        unc = PC.PromptEmbeddings(None, torch.zeros(1, 4), None)          # class_labels row: 0 = uncond, 1 = cond
        txt = PC.PromptEmbeddings(None, torch.ones(1, 4), None)
        mode = getattr(PC.PCStreamChoice, cs["mode"])
        _, x0_pred = PC.forward_directional(ldm, xt, t, lat, unc, txt, 3.0, eta=1)
        rec = []
        orig = PC.forward_directional

        def spy(*a, **k):
            r = orig(*a, **k)
            rec.append((k["eigvecs"].clone(), r[1].clone()))
            return r
        PC.forward_directional = spy
        try:
            torch.manual_seed(1000 + cs["seed"])
            eigvecs, eigval, in_corr, in_norm, interm_vec, interm_val = PC.get_eigenvectors(
                ldm, xt, txt, unc, lat, mask, t, x0_pred, mode, 1e-3, 3.0, cs["iters"], False, 1, cs["n_ev"])
        finally:
            PC.forward_directional = orig
        n = cs["n_ev"]
        k = cs["name"]
        scaled_in = torch.stack([r[0].reshape(n, D) for r in rec])            # perturbation fed to iteration i
        x0p = torch.stack([r[1].reshape(n, D) for r in rec])                  # x0_pred of the perturbed rows
        out.update({f"{k}_n_ev": n, f"{k}_iters": cs["iters"], f"{k}_mode": mode.value, f"{k}_t": int(t),
                    f"{k}_xt": xt, f"{k}_lat": lat, f"{k}_mask": mask, f"{k}_x0_pred": x0_pred,
                    f"{k}_U": fac[0], f"{k}_lam_u": fac[1], f"{k}_lam_c": fac[2],
                    f"{k}_scaled_in": scaled_in, f"{k}_x0p": x0p, f"{k}_eigvecs": eigvecs.reshape(n, D),
                    f"{k}_eigval": eigval.reshape(-1), f"{k}_in_norm": torch.stack([v.reshape(-1) for v in in_norm]),
                    f"{k}_in_corr": torch.stack([v.reshape(-1) for v in in_corr]),
                    f"{k}_interm_keys": np.asarray(sorted(interm_vec), np.int64)})
        for i in interm_vec:
            out[f"{k}_interm_vec_{i}"] = interm_vec[i].reshape(n, D)
            out[f"{k}_interm_val_{i}"] = interm_val[i].reshape(-1)
        if k == "a":
            # forward_directional with a shifted input, all three stream choices (pc_drift.py:41-93)
            ev = torch.nn.functional.normalize(torch.randn(2, C_, H_, W_, generator=g).reshape(2, -1), dim=1
                                               ).reshape(2, C_, H_, W_)
            xt2 = torch.randn(2, C_, H_, W_, generator=g)
            lat2 = torch.randn(2, C_, H_, W_, generator=g)
            for md in ("BOTH", "TEXT", "UNCOND"):
                pr, x0 = PC.forward_directional(ldm, xt2, t, lat2, unc, txt, 3.0, eta=1, eigvecs=ev, amount=0.7,
                                                mode=getattr(PC.PCStreamChoice, md))
                out[f"fd_prev_{md}"], out[f"fd_x0_{md}"] = pr, x0
            out["fd_xt"], out["fd_lat"], out["fd_ev"] = xt2, lat2, ev
            # apply_drift (pc_drift.py:201-278) on the extracted directions
            eigdata = {int(t): dict(eigvec=eigvecs, eigval=eigval, interm_eigvecs=interm_vec, interm_eigvals=interm_val)}
            xm1, x0p1 = PC.forward_directional(ldm, xt, t, lat, unc, txt, 3.0, eta=1)
            for eta in (1, 0):
                for shifted in (True, False):
                    o = PC.apply_drift(ldm, xm1, x0p1, t, ldm.model.scheduler.timesteps, N, eigdata, lat,
                                       torch.device("cpu"), use_shifted_x0_for_noisepred=shifted, amount=1.5, eta=eta,
                                       ev_nums=[1, 3])
                    out[f"ad_eta{eta}_sh{int(shifted)}"] = o
            out["ad_xm1"], out["ad_x0p"] = xm1, x0p1
    save("pc_drift.npz", n_steps=N, shape=np.asarray([C_, H_, W_]), **out)


def golden_pc_unet():
    """Reference get_eigenvectors (unmodified, const = 1e-3, fp32 CPU) through the vendored tiny U-Net: the finite
    differences it forms per iteration (perturbation in, posterior mean out) — the yardstick for the device path's
    Jacobian-vector products through the real network (tests/test_gpu_pc_drift.py::test_unet_jvp_resolves_with_fd_const)."""
    ref = ref_import.load()
    PC = ref.pc_drift
    cfg = C.preset("tiny-audioldm")
    w = U.synthetic_weights(cfg, seed=0)
    N, n_ev, iters = 20, 2, 6
    model = make_fake_wrapper(ref, cfg, w, N)
    g = torch.Generator().manual_seed(81)
    xt = torch.randn(1, 8, 16, 16, generator=g)
    lat = torch.randn(1, 8, 16, 16, generator=g)
    mask = torch.ones(1, 8, 16, 16)
    t = model.model.scheduler.timesteps[8]
    unc = PC.PromptEmbeddings(None, prompt_vector(""), None)
    txt = PC.PromptEmbeddings(None, prompt_vector("a dog barking"), None)
    with torch.no_grad():
        _, x0_pred = PC.forward_directional(model, xt, t, lat, unc, txt, 3.0, eta=1)
    rec = []
    orig = PC.forward_directional

    def spy(*a, **k):
        r = orig(*a, **k)
        rec.append((k["eigvecs"].clone(), r[1].clone()))
        return r
    PC.forward_directional = spy
    try:
        torch.manual_seed(82)
        PC.get_eigenvectors(model, xt, txt, unc, lat, mask, t, x0_pred, PC.PCStreamChoice.BOTH, 1e-3, 3.0, iters, False, 1,
                            n_ev)
    finally:
        PC.forward_directional = orig
    save("pc_unet.npz", n_steps=N, t=int(t), xt=xt, lat=lat, x0_pred=x0_pred, uncond=prompt_vector(""),
         cond=prompt_vector("a dog barking"), scaled_in=torch.stack([r[0] for r in rec]), x0p=torch.stack([r[1] for r in rec]))


def golden_sdedit():
    """SDEdit flow of code/main_run_sdedit.py:78-100 (pre-drawn latents, add_noise at timesteps[skip], forward_directional
    loop), reference pc_drift.forward_directional unmodified on the fake AudioLDM wrapper (vendored UNetModel); the
    diffusers scheduler's add_noise / step / init_noise_sigma come from the MiniDDIM restatement."""
    ref = ref_import.load()
    PC = ref.pc_drift
    cfg = C.preset("tiny-audioldm")
    w = U.synthetic_weights(cfg, seed=0)
    N, tstart = 10, 6
    model = make_fake_wrapper(ref, cfg, w, N)
    g = torch.Generator().manual_seed(71)
    w0 = 0.5 * torch.randn(1, 8, 16, 16, generator=g)
    timesteps = model.model.scheduler.timesteps
    latents = [torch.randn(1, 8, 16, 16, generator=g) * model.model.scheduler.init_noise_sigma
               for _ in range(len(timesteps) + 1)]                                                   # :79-87
    skip = N - tstart                                                                                # :89
    noise = torch.randn(w0.shape, generator=g)
    xt = model.model.scheduler.add_noise(w0, noise, timesteps[skip].unsqueeze(0))                    # :92-93
    unc = PC.PromptEmbeddings(None, prompt_vector(""), None)
    txt = PC.PromptEmbeddings(None, prompt_vector("a cat"), None)
    x_start = xt.clone()
    with torch.no_grad():
        for it, t in enumerate(timesteps[skip:]):                                                    # :97-100
            xt, _ = PC.forward_directional(model, xt, t, latents[skip + it + 1], unc, txt, 5.0, eta=1)
    save("sdedit.npz", n_steps=N, tstart=tstart, w0=w0, noise=noise, latents=torch.cat(latents), x_start=x_start,
         w_edit=xt, uncond=prompt_vector(""), tgt=prompt_vector("a cat"), cfg_tar=5.0)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "pc"):
        golden_pc_drift()
    if what in ("all", "sdedit"):
        golden_sdedit()
    if what in ("all", "pcunet"):
        golden_pc_unet()
    if what in ("all", "ddim"):
        golden_ddim()
    if what in ("all", "loops"):
        main()
    if what in ("all", "stft"):
        golden_stft()
    if what in ("all", "ends"):
        golden_ends()
