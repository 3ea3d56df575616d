"""GPU parity of the audio ends against golden outputs of the reference's vendored
modules (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def test_stft_mel_vs_reference_tacotron():
    """audioldm/audio/stft.py TacotronSTFT.mel_spectrogram (dense-DFT conv1d on the CPU) vs ae_stft_mel.
    Tolerance: 2e-3 abs on the log-mel (fp32 direct DFT vs fp32 conv accumulation), 1e-3 abs on magnitudes."""
    from audioeditingcode_b200.audio import TacotronSTFT
    g = load_golden("stft_mel.npz")
    fn = TacotronSTFT(1024, 160, 1024, 64, 16000, 0, 8000, device="cuda")
    assert torch.equal(fn.mel_basis.cpu(), g["mel_basis"])
    assert torch.equal(fn.window.cpu(), g["window"])
    mel, logmag, energy = fn.mel_spectrogram(g["wav"][None].cuda())
    assert mel.shape == (1, 64, 201)
    assert (mel[0].cpu() - g["mel"]).abs().max().item() < 2e-3
    assert (logmag[0].cpu().exp() - g["logmag"].exp()).abs().max().item() < 1e-3
    assert torch.allclose(energy[0].cpu(), g["energy"], rtol=1e-4, atol=1e-3)


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()


def test_vae_encode_decode_vs_vendored_modules():
    """VAEEngine (tcgen05 convs, GN kernels, unfused single-head attention) vs golden outputs of the reference's
    vendored Encoder / Decoder (variational_autoencoder/modules.py) with identical synthetic weights.
    Tolerance: bf16 operands, fp32 accumulation / residual stream -> rel-L2 <= max(1e-2, error of stock PyTorch
    bf16 autocast on the same network, stored in the fixture)."""
    from audioeditingcode_b200.ends import VAEEngine, vae_weight_shapes, synthetic, VAE_SCALING
    g = load_golden("vae_ends.npz")
    vae = VAEEngine("cuda", synthetic(vae_weight_shapes(), 0), VAE_SCALING)
    z = vae.encode_mode(g["x"].cuda())
    assert z.shape == g["z"].shape
    assert _rel(z, g["z"]) < max(1e-2, float(g["bf16_autocast_err_encode"]))
    mom = vae.encode_moments(g["x"].cuda())
    assert _rel(mom, g["moments"]) < max(1e-2, float(g["bf16_autocast_err_encode"]))
    dec = vae.decode(g["z"].cuda())
    assert dec.shape == g["decoded"].shape
    r = _rel(dec, g["decoded"])
    print(f"vae decode rel-L2 {r:.2e} (torch-bf16 {float(g['bf16_autocast_err_decode']):.2e})")
    assert r < max(1e-2, float(g["bf16_autocast_err_decode"]))


def test_hifigan_vs_vendored_generator():
    """HiFiGANEngine (1-D dilated implicit-GEMM convs, phase-decomposed transposed convs) vs the golden waveform of
    the reference's vendored Generator (hifigan/models.py).  Tolerance rel-L2 <= 2e-2 (15 residual MRF blocks deep)."""
    from audioeditingcode_b200.ends import HiFiGANEngine, hifigan_weight_shapes, synthetic
    g = load_golden("hifigan_ends.npz")
    voc = HiFiGANEngine("cuda", synthetic(hifigan_weight_shapes(), 0))
    wav = voc(g["mel"][0].cuda())
    assert wav.shape == g["wav"][0].shape
    r = _rel(wav, g["wav"][0])
    print(f"hifigan rel-L2 {r:.2e} (torch-bf16 {float(g['bf16_autocast_err']):.2e})")
    assert r < 2e-2
    assert (wav.cpu() - g["wav"][0]).abs().max().item() < 0.05 * g["wav"].abs().max().item() + 1e-4
    # batch axis (the edited / original pair of main_run.py:184-185 in one launch sequence): each clip's waveform equals
    # its single-clip result (per-clip zero padding; a different M may pick another tile / split-K plan, so the
    # comparison is to fp32 summation order, not bits)
    mel2 = torch.stack([g["mel"][0], g["mel"][0].flip(0) * 0.5 - 1.0]).cuda()
    w2 = voc(mel2)
    assert w2.shape == (2, wav.shape[0])
    r0, r1 = _rel(w2[0], wav), _rel(w2[1], voc(mel2[1]))
    print(f"batched vocoder vs single calls: rel-L2 {r0:.2e} {r1:.2e}")
    from audioeditingcode_b200 import _lib
    tol = 8e-3 if _lib.load().ae_operand_dtype() == 0 else 2e-3          # bf16-operand build: 8x the operand rounding
    assert r0 < tol and r1 < tol
    assert _rel(w2[0], g["wav"][0]) < 2e-2


def test_wrapper_ends_roundtrip_shapes():
    """models.py wrapper methods vae_encode (front-pads T to a multiple of 4, models.py:497-498) / vae_decode /
    decode_to_mel keep the reference's shapes."""
    from audioeditingcode_b200 import models
    m = models.load_model("synthetic/audioldm-tiny", torch.device("cuda"), 10)
    x = torch.randn(1, 1, 126, 64, device="cuda") * 2 - 4
    w0 = m.vae_encode(x)
    assert w0.shape == (1, 8, 32, 16) and w0.dtype == torch.float32
    xd = m.vae_decode(w0)
    assert xd.shape == (1, 1, 128, 64)
    wav = m.decode_to_mel(xd)
    assert wav.dim() == 2 and wav.shape[0] == 1 and abs(wav.shape[1] - 128 * 160) <= 64


def test_tango_posterior_sample_encode():
    """TangoWrapper.vae_encode = get_first_stage_encoding(encode_first_stage(x)) = posterior.sample() * scale_factor
    (models.py:439-447; DiagonalGaussianDistribution, distributions.py:24-73: logvar clamped to [-30, 20],
    sample = mean + std * randn): with the generator seeded identically the sample is mean + exp(logvar / 2) * eps of the
    moments the vendored-Encoder golden pins, and its front padding / length limit behave like the reference's."""
    from audioeditingcode_b200 import models, unet_config as C
    g = load_golden("vae_ends.npz")
    m = models.load_model("synthetic/tango-tiny", torch.device("cuda"), 10, config=C.preset("tiny-tango"))
    x = g["x"].cuda()
    ends = m._ends()
    mom = ends.vae().encode_moments(x)
    mean, logvar = mom[:, :8], torch.clamp(mom[:, 8:], -30.0, 20.0)
    torch.manual_seed(5)
    z = m.vae_encode(x)
    torch.manual_seed(5)
    eps = torch.randn_like(mean)
    want = (mean + torch.exp(0.5 * logvar) * eps) * ends.vae().scaling
    assert torch.allclose(z, want, atol=1e-6, rtol=1e-6)
    torch.manual_seed(6)
    assert not torch.equal(m.vae_encode(x), z)
    with pytest.raises(RuntimeWarning):                       # models.py:444-445
        m.vae_encode(torch.zeros(1, 1, 1704, 64, device="cuda"))
    xp = m.vae_encode(x[:, :, :x.shape[2] - 2])               # T % 4 != 0 -> front-padded to a multiple of 4 (:441-442)
    assert xp.shape[2] == (x.shape[2] - 2 + 3) // 4
