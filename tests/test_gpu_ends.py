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
