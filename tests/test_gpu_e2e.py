"""The call sequence of the reference's CLI, EXECUTED end to end on the CUDA path through the drop-in modules: wav file ->
load_audio (mel-STFT) -> vae_encode -> inversion_forward_process -> inversion_reverse_process (or the DDIM branch) ->
vae_decode -> decode_to_mel (twice: edited and original) -> wav file.  This is what code/main_run.py:133-225 does between
argument parsing and wandb / matplotlib (absent from this image, as is /root/reference on the GPU box — so the script
itself cannot be run; tests/test_dropin_surface.py checks its call surface statically).  Tiny seeded checkpoints of the
three wrapper families; asserts shapes, finiteness, determinism (same seed -> same bits) and that the edit acted."""
import importlib
import os
import sys
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("models", "utils", "pc_drift", "ddm_inversion", "ddm_inversion.inversion_utils", "ddm_inversion.ddim_inversion")


class _Dropin:
    """`import models` etc. resolve to dropin/ inside the block, and the shims are unregistered afterwards."""

    def __enter__(self):
        self.path = list(sys.path)
        self.saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in NAMES}
        sys.path[:0] = [os.path.join(ROOT, "dropin"), ROOT]
        return self

    def __exit__(self, *exc):
        sys.path[:] = self.path
        for k in [k for k in list(sys.modules) if k in NAMES]:
            del sys.modules[k]
        sys.modules.update(self.saved)


@pytest.fixture(autouse=True)
def _restore_torch_flags():
    """set_reproducability (utils.py:113-116) flips process-wide torch flags (TF32 'high' matmuls): keep them from leaking
    into the tests that run after this module."""
    prec, tf32 = torch.get_float32_matmul_precision(), torch.backends.cudnn.allow_tf32
    yield
    torch.set_float32_matmul_precision(prec)
    torch.backends.cudnn.allow_tf32 = tf32


def _write_clip(path, seconds=1.3, sr=16000):
    t = np.arange(int(seconds * sr)) / sr
    rng = np.random.default_rng(5)
    x = 0.4 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 1330 * t + 1.0) + 0.05 * rng.standard_normal(t.size)
    pcm = np.clip(x * 32768, -32768, 32767).astype("<i2")
    with wave.open(str(path), "w") as f:
        f.setnchannels(1), f.setsampwidth(2), f.setframerate(sr)
        f.writeframes(pcm.tobytes())


def _edit(tmp_path, model_id, preset, mode, seed, tag):
    with _Dropin():
        load_model = importlib.import_module("models").load_model
        U = importlib.import_module("utils")
        IU = importlib.import_module("ddm_inversion.inversion_utils")
        DI = importlib.import_module("ddm_inversion.ddim_inversion")
        from audioeditingcode_b200 import unet_config as C, audio as A
        N, tstart = 20, torch.tensor([12], dtype=torch.int)
        skip = N - tstart
        U.set_reproducability(seed, extreme=False)                                   # main_run.py:71
        ldm_stable = load_model(model_id, "cuda:0", N, config=C.preset(preset))      # :133
        clip = tmp_path / "clip.wav"
        if not clip.exists():
            _write_clip(clip)
        x0, sr, duration = U.load_audio(str(clip), ldm_stable.get_fn_STFT(), device="cuda:0", stft=True,
                                        model_sr=ldm_stable.get_sr())               # :134-135
        assert x0.dim() == 4 and x0.shape[1] == 1 and x0.shape[3] == 64 and abs(duration - 1.3) < 1e-3
        with torch.inference_mode():
            w0 = ldm_stable.vae_encode(x0)                                            # :138
            if mode == "ddim":
                wT = DI.ddim_inversion(ldm_stable, w0, ["a tone"], 3.0, num_inference_steps=N, skip=skip[0])
                w_edit = DI.text2image_ldm_stable(ldm_stable, ["a bell"], N, 5.0, wT, skip=skip)
            else:
                wt, zs, wts, extra_info = IU.inversion_forward_process(
                    ldm_stable, w0, etas=1.0, prompts=["a tone"], cfg_scales=[3.0], prog_bar=True, num_inference_steps=N,
                    cutoff_points=None, numerical_fix=True, duration=duration)      # :147-154
                w_edit, _ = IU.inversion_reverse_process(
                    ldm_stable, xT=wts, tstart=tstart, fix_alpha=0.1, etas=1.0, prompts=["a bell"], neg_prompts=[""],
                    cfg_scales=[5.0], prog_bar=True, zs=zs[:int(N - min(skip))], cutoff_points=None, duration=duration,
                    extra_info=extra_info)                                           # :168-181
            x0_dec = ldm_stable.vae_decode(w_edit)                                    # :201
            if x0_dec.dim() < 4:
                x0_dec = x0_dec[None, :, :, :]
            audio = ldm_stable.decode_to_mel(x0_dec)                                  # :207-208
            orig_audio = ldm_stable.decode_to_mel(x0)
        out, orig = tmp_path / f"edit_{tag}.wav", tmp_path / f"orig_{tag}.wav"
        A.save_wav(str(out), audio, sample_rate=sr)                                   # torchaudio.save, :223-224
        A.save_wav(str(orig), orig_audio, sample_rate=sr)
        pcm, sr2 = A._load_wav(str(out))
        pcm_o, _ = A._load_wav(str(orig))
    assert sr2 == sr == 16000
    return w0.clone(), w_edit.clone(), x0_dec.clone(), pcm, pcm_o


@pytest.mark.parametrize("model_id,preset", [("synthetic/audioldm2-tiny", "tiny-audioldm2"),
                                             ("synthetic/audioldm-tiny", "tiny-audioldm"),
                                             ("synthetic/tango-tiny", "tiny-tango")])
@pytest.mark.parametrize("mode", ["ours", "ddim"])
def test_main_run_sequence_wav_to_wav(tmp_path, model_id, preset, mode):
    w0, w_edit, x0_dec, pcm, pcm_o = _edit(tmp_path, model_id, preset, mode, seed=3, tag="a")
    assert w_edit.shape == w0.shape and w0.shape[1] == 8 and w0.shape[3] == 16
    assert torch.isfinite(w_edit).all() and torch.isfinite(x0_dec).all()
    assert x0_dec.shape[1] == 1 and x0_dec.shape[3] == 64 and x0_dec.shape[2] == 4 * w0.shape[2]
    n_expected = x0_dec.shape[2] * 160
    assert pcm.dim() == 2 and pcm.shape[0] == 1 and abs(pcm.shape[1] - n_expected) <= 64 and pcm.abs().max() <= 1.0
    assert pcm_o.shape[1] >= 126 * 160 and float(pcm.abs().max()) > 0
    assert (w_edit - w0).abs().max().item() > 1e-3                       # the target prompt / guidance moved the latent
    # same seed -> the same bits, end to end (TANGO draws its VAE posterior sample and everything else from the seed)
    w0b, w_editb, _, pcmb, _ = _edit(tmp_path, model_id, preset, mode, seed=3, tag="b")
    assert torch.equal(w0, w0b) and torch.equal(w_edit, w_editb) and torch.equal(pcm, pcmb)
