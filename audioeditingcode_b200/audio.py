"""Audio front end — drop-in for the pieces of code/audioldm/audio/{stft,tools,audio_processing}.py and
code/audioldm/utils.py that the editing path calls (`TacotronSTFT`, `wav_to_fbank`,
`get_duration`):

    TacotronSTFT(filter_length, hop_length, win_length, n_mel_channels, sampling_rate, mel_fmin, mel_fmax)
        .mel_spectrogram(y) -> (mel [B, n_mels, T], log_magnitudes [B, n_fft/2+1, T], energy [B, T])   stft.py:159-180
    wav_to_fbank(filename, target_length, fn_STFT) -> (fbank [T, 64], log_magnitudes [T, 513], waveform)  tools.py:67-85
    read_wav_file / normalize_wav / pad_wav / _pad_spec                                                 tools.py:18-64
    get_duration(fname)                                                                                 audioldm/utils.py:17-21

The reference evaluates the STFT as a dense windowed-DFT conv1d and moves the result to the CPU (stft.py:67-72);
here the whole log-mel computation is one device kernel (ae_stft_mel) and stays on the device.
"""
from __future__ import annotations

import contextlib
import wave
from typing import Optional

import numpy as np
import torch

from . import _lib
from .models import _ptr, _stream


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney'),
    restated from the published algorithm (librosa is not a dependency of this package) — stft.py:141-143."""
    def hz_to_mel(f):
        f = np.asanyarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asanyarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


class TacotronSTFT(torch.nn.Module):
    def __init__(self, filter_length, hop_length, win_length, n_mel_channels, sampling_rate, mel_fmin, mel_fmax,
                 device: Optional[torch.device] = None):
        super().__init__()
        assert filter_length >= win_length
        self.filter_length = filter_length
        self.hop_length = hop_length
        self.win_length = win_length
        self.n_mel_channels = n_mel_channels
        self.sampling_rate = sampling_rate
        self.dev = torch.device(device) if device is not None else torch.device("cuda")
        # scipy.signal.get_window('hann', win_length, fftbins=True) (periodic Hann), centre-padded to filter_length
        n = np.arange(win_length, dtype=np.float64)
        win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
        lpad = (filter_length - win_length) // 2
        win = np.pad(win, (lpad, filter_length - win_length - lpad))
        self.register_buffer("window", torch.from_numpy(win).float())
        self.register_buffer("mel_basis", torch.from_numpy(
            slaney_mel_basis(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)))

    def mel_spectrogram(self, y: torch.Tensor, normalize_fun=torch.log):
        """y: [B, T] in [-1, 1].  Returns (mel [B, n_mels, frames], log_magnitudes [B, bins, frames], energy)."""
        assert torch.min(y.data) >= -1, torch.min(y.data)
        assert torch.max(y.data) <= 1, torch.max(y.data)
        if normalize_fun is not torch.log:
            raise NotImplementedError("only the default torch.log compression is used on the editing path")
        lib = _lib.load()
        y = y.to(self.dev, torch.float32).contiguous()
        window = self.window.to(self.dev)
        mel_basis = self.mel_basis.to(self.dev)
        B, n = y.shape
        frames = n // self.hop_length + 1
        bins = self.filter_length // 2 + 1
        mel = torch.empty(B, frames, self.n_mel_channels, device=self.dev)
        mag = torch.empty(B, frames, bins, device=self.dev)
        for b in range(B):
            _lib.check(lib.ae_stft_mel(_ptr(y[b]), n, self.filter_length, self.hop_length, _ptr(window), _ptr(mel_basis),
                                       self.n_mel_channels, frames, _ptr(mag[b]), _ptr(mel[b]), _stream()), "ae_stft_mel")
        mag_t = mag.transpose(1, 2)
        log_magnitudes = torch.log(torch.clamp(mag_t, min=1e-5))
        energy = torch.norm(mag_t, dim=1)
        return mel.transpose(1, 2), log_magnitudes, energy


# ------------------------------------------------------------------------------------------------- tools.py
def get_duration(fname: str) -> float:                                           # audioldm/utils.py:17-21
    with contextlib.closing(wave.open(fname, "r")) as f:
        return f.getnframes() / float(f.getframerate())


def _pad_spec(fbank: torch.Tensor, target_length: int = 1024) -> torch.Tensor:   # tools.py:18-31
    n_frames = fbank.shape[0]
    p = target_length - n_frames
    if p > 0:
        fbank = torch.nn.functional.pad(fbank, (0, 0, 0, p))
    elif p < 0:
        fbank = fbank[0:target_length, :]
    if fbank.size(-1) % 2 != 0:
        fbank = fbank[..., :-1]
    return fbank


def pad_wav(waveform: np.ndarray, segment_length: Optional[int]) -> np.ndarray:    # tools.py:34-44
    waveform_length = waveform.shape[-1]
    assert waveform_length > 100, "Waveform is too short, %s" % waveform_length
    if segment_length is None or waveform_length == segment_length:
        return waveform
    elif waveform_length > segment_length:
        return waveform[:segment_length]
    temp_wav = np.zeros((1, segment_length))
    temp_wav[:, :waveform_length] = waveform
    return temp_wav


def normalize_wav(waveform: np.ndarray) -> np.ndarray:                             # tools.py:46-49
    waveform = waveform - np.mean(waveform)
    waveform = waveform / (np.max(np.abs(waveform)) + 1e-8)
    return waveform * 0.5


def _load_wav(filename: str):
    """16/32-bit PCM RIFF reader on the stdlib (torchaudio.load needs an I/O backend this image may lack)."""
    with contextlib.closing(wave.open(filename, "r")) as f:
        sr, nch, width, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
        raw = f.readframes(n)
    if width == 2:
        data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported sample width {width}")
    return torch.from_numpy(data.reshape(-1, nch).T.copy()), sr


def read_wav_file(filename: str, segment_length: int) -> np.ndarray:                # tools.py:52-64
    waveform, sr = _load_wav(filename)
    if sr != 16000:
        import torchaudio
        waveform = torchaudio.functional.resample(waveform, orig_freq=sr, new_freq=16000)
    waveform = waveform.numpy()[0, ...]
    waveform = normalize_wav(waveform)
    waveform = waveform[None, ...]
    waveform = pad_wav(waveform, segment_length)
    waveform = waveform / np.max(np.abs(waveform))
    waveform = 0.5 * waveform
    return waveform


def get_mel_from_wav(audio, _stft: TacotronSTFT):                                   # tools.py:6-15 (stays on device)
    audio = torch.clip(torch.as_tensor(audio, dtype=torch.float32).unsqueeze(0), -1, 1)
    melspec, log_magnitudes_stft, energy = _stft.mel_spectrogram(audio)
    return melspec.squeeze(0), log_magnitudes_stft.squeeze(0), energy.squeeze(0)


def wav_to_fbank(filename: str, target_length: int = 1024, fn_STFT: Optional[TacotronSTFT] = None):   # tools.py:67-85
    assert fn_STFT is not None
    waveform = read_wav_file(filename, target_length * 160)
    waveform = torch.FloatTensor(waveform[0, ...])
    fbank, log_magnitudes_stft, energy = get_mel_from_wav(waveform, fn_STFT)
    fbank = fbank.T
    log_magnitudes_stft = log_magnitudes_stft.T
    fbank, log_magnitudes_stft = _pad_spec(fbank, target_length), _pad_spec(log_magnitudes_stft, target_length)
    return fbank, log_magnitudes_stft, waveform


def save_wav(path: str, wav, sample_rate: int = 16000) -> None:
    """16-bit PCM RIFF writer on the stdlib (what main_run.py:223-224 asks of torchaudio.save, which needs an I/O
    backend this image may lack).  wav: float tensor in [-1, 1] ([T] or [1, T]; converted on the device when it lives
    there, ae_wave_to_int16) or an int16 numpy array as TangoWrapper.decode_to_mel returns."""
    if torch.is_tensor(wav):
        w = wav.detach().reshape(-1).float().contiguous()
        if w.is_cuda:
            pcm = torch.empty(w.shape, dtype=torch.int16, device=w.device)
            lib = _lib.load()
            _lib.check(lib.ae_wave_to_int16(_ptr(w), w.numel(), _ptr(pcm), _stream()), "ae_wave_to_int16")
            data = pcm.cpu().numpy()
        else:
            data = np.clip(w.numpy() * 32768.0, -32768.0, 32767.0).astype(np.int16)
    else:
        data = np.asarray(wav).reshape(-1).astype(np.int16)
    with contextlib.closing(wave.open(path, "w")) as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(data.astype("<i2").tobytes())
