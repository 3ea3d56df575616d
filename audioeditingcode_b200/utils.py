"""Helpers of the editing scripts — drop-in for the pieces of code/utils.py on the audio path:
load_audio (:53-95), get_spec (:49-50), set_reproducability (:98-116), get_height_of_spectrogram (:119-135),
get_text_embeddings (:217-231).  Plotting / wandb helpers (:141-214) and load_image (:16-46) are out of scope."""
from __future__ import annotations

import os
import random
from typing import List, Optional, Tuple

import numpy as np
import torch

from .models import PipelineWrapper
from .pc_drift import PromptEmbeddings
from . import audio as _audio


def get_spec(wav: torch.Tensor, fn_STFT: torch.nn.Module) -> torch.Tensor:
    return fn_STFT.mel_spectrogram(torch.clip(wav[:, 0], -1, 1))[0]


def load_audio(audio_path, fn_STFT, left: int = 0, right: int = 0, device: Optional[torch.device] = None,
               return_wav: bool = False, stft: bool = False, model_sr: Optional[int] = None):
    if stft:  # AudioLDM/tango loading to spectrogram
        if type(audio_path) is str:
            duration = _audio.get_duration(audio_path)
            mel, _, wav = _audio.wav_to_fbank(audio_path, target_length=int(duration * 102.4), fn_STFT=fn_STFT)
            mel = mel.unsqueeze(0)
        else:
            mel = audio_path
        c, h, w = mel.shape
        left = min(left, w - 1)
        right = min(right, w - left - 1)
        mel = mel[:, :, left:w - right]
        mel = mel.unsqueeze(0).to(device)
        if return_wav:
            return mel, 16000, duration, wav
        return mel, model_sr, duration
    waveform, sr = _audio._load_wav(audio_path)
    if sr != model_sr:
        import torchaudio
        waveform = torchaudio.functional.resample(waveform, orig_freq=sr, new_freq=model_sr)
    waveform = waveform - torch.mean(waveform)
    waveform = waveform / (torch.max(torch.abs(waveform)) + 1e-8) * 0.5
    duration = waveform.shape[-1] / model_sr
    return torch.FloatTensor(waveform), model_sr, duration


def set_reproducability(seed: int, extreme: bool = True) -> None:
    if seed is not None:
        torch.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
        random.seed(seed)
        np.random.seed(seed)
        if extreme:
            torch.use_deterministic_algorithms(True)
            os.environ["CUBLAS_WORKSPACE_CONFIG"] = ":4096:8"
        torch.backends.cudnn.benchmark = False
    # libaedit's kernels are deterministic and TF32-free by construction; the torch flags below only matter for
    # the torch-side plumbing and are kept for script compatibility (utils.py:113-116)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("high")      # utils.py:116 (torch-side matmuls only, e.g. the text encoders)


def get_height_of_spectrogram(length: int, ldm_stable: PipelineWrapper) -> int:
    vocoder_upsample_factor = np.prod(ldm_stable.model.vocoder.config.upsample_rates) / \
        ldm_stable.model.vocoder.config.sampling_rate
    if length is None:
        length = ldm_stable.model.unet.config.sample_size * ldm_stable.model.vae_scale_factor * vocoder_upsample_factor
    height = int(length / vocoder_upsample_factor)
    if height % ldm_stable.model.vae_scale_factor != 0:
        height = int(np.ceil(height / ldm_stable.model.vae_scale_factor)) * ldm_stable.model.vae_scale_factor
        print(f"Audio length in seconds {length} is increased to {height * vocoder_upsample_factor} "
              f"so that it can be handled by the model. It will be cut to {length} after the denoising process.")
    return height


def get_text_embeddings(target_prompt: List[str], target_neg_prompt: List[str], ldm_stable: PipelineWrapper
                        ) -> Tuple[torch.Tensor, PromptEmbeddings, PromptEmbeddings]:
    text_hs, text_cl, text_mask = ldm_stable.encode_text(target_prompt)
    un_hs, un_cl, un_mask = ldm_stable.encode_text(target_neg_prompt)
    text_emb = PromptEmbeddings(embedding_hidden_states=text_hs, boolean_prompt_mask=text_mask,
                                embedding_class_lables=text_cl)
    uncond_emb = PromptEmbeddings(embedding_hidden_states=un_hs, boolean_prompt_mask=un_mask,
                                  embedding_class_lables=un_cl)
    return text_cl, text_emb, uncond_emb
