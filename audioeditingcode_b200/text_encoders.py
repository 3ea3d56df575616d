"""Text conditioning stage: prompt strings -> the frozen conditioning tensors the
U-Net kernels consume, once per prompt per run, OUTSIDE the per-step loop.

Reference behaviour reproduced (code/models.py):
    AudioLDMWrapper.encode_text   :511-537  RoBERTa tokenizer (padding='max_length') -> ClapTextModelWithProjection
                                            text_embeds -> L2 normalise -> (None, [P,512], None)
    AudioLDM2Wrapper.encode_text  :599-677  CLAP get_text_features ([P,1,512], mask of ones) + T5 encoder ([P,L,1024],
                                            padding=True) -> projection model (Linear + learned SOS/EOS per stream,
                                            concatenated) -> GPT-2 generates 8 continuous tokens ->
                                            (generated [P,8,768], T5 hidden [P,L,1024], T5 mask [P,L])
    TangoWrapper.encode_text      :455-460  tango AudioDiffusion.encode_text: T5 tokenizer (padding=True) -> T5 encoder
                                            -> ([P,L,1024], None, bool mask [P,L])

As SURVEY.md §8(a13) prescribes, the encoders themselves stay in `transformers` (frozen inputs of the path, not
per-step work): they are loaded from a LOCAL checkpoint directory laid out like the diffusers pipelines the reference
loads (`tokenizer/`, `text_encoder/`, `tokenizer_2/`, `text_encoder_2/`, `projection_model/`, `language_model/`), or a
TANGO snapshot (`main_config.json`, `pytorch_model_main.bin` with `text_encoder.*` keys).  The two pieces that live in
diffusers — `AudioLDM2ProjectionModel` and `AudioLDM2Pipeline.generate_language_model` — are restated here from the
published algorithm ([UPSTREAM], diffusers is not installed): projection = Linear per stream, learned SOS / EOS
embeddings around each stream, concatenation; generation = 8 steps of GPT2Model on `inputs_embeds`, each appending
the last hidden state (no sampling, no logits).

Embeddings can be cached on disk per (checkpoint, family, prompt) with AEDIT_TEXT_CACHE=<dir> (the encoders are
deterministic): a dataset run over many clips with the same prompts then never touches the encoders again.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Dict, List, Optional, Tuple

import torch


def _load_state(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def _find(dirname: str, *names: str) -> Optional[str]:
    for n in names:
        p = os.path.join(dirname, n)
        if os.path.exists(p):
            return p
    return None


class _DiskCache:
    def __init__(self, tag: str):
        self.dir = os.environ.get("AEDIT_TEXT_CACHE")
        self.tag = tag

    def _path(self, prompt: str) -> str:
        h = hashlib.sha1((self.tag + "\0" + prompt).encode()).hexdigest()
        return os.path.join(self.dir, h + ".pt")

    def get(self, prompt: str):
        if not self.dir:
            return None
        p = self._path(prompt)
        return torch.load(p, map_location="cpu", weights_only=True) if os.path.exists(p) else None

    def put(self, prompt: str, value) -> None:
        if not self.dir:
            return
        os.makedirs(self.dir, exist_ok=True)
        tmp = self._path(prompt) + f".tmp{os.getpid()}"
        torch.save(value, tmp)
        os.replace(tmp, self._path(prompt))


def _tokenize(tokenizer, prompts: List[str], max_length_padding: bool):
    """The reference's tokenizer call (models.py:512-518, :607-613) including its truncation notice (:523-528)."""
    text_inputs = tokenizer(prompts, padding="max_length" if max_length_padding else True,
                            max_length=tokenizer.model_max_length, truncation=True, return_tensors="pt")
    untruncated_ids = tokenizer(prompts, padding="longest", return_tensors="pt").input_ids
    ids = text_inputs.input_ids
    if untruncated_ids.shape[-1] >= ids.shape[-1] and not torch.equal(ids, untruncated_ids):
        removed = tokenizer.batch_decode(untruncated_ids[:, tokenizer.model_max_length - 1: -1])
        print("The following part of your input was truncated because the text encoder can only handle sequences up to"
              f" {tokenizer.model_max_length} tokens: {removed}")
    return ids, text_inputs.attention_mask


class ClapTextEncoder:
    """AudioLDM-1 conditioning (models.py:511-537)."""

    def __init__(self, ckpt_dir: str, device):
        from transformers import AutoTokenizer, ClapTextModelWithProjection
        self.device = torch.device(device)
        self.tokenizer = AutoTokenizer.from_pretrained(os.path.join(ckpt_dir, "tokenizer"), local_files_only=True)
        self.text_encoder = ClapTextModelWithProjection.from_pretrained(os.path.join(ckpt_dir, "text_encoder"),
                                                                         local_files_only=True).to(self.device).eval()
        self._cache = _DiskCache(f"audioldm:{os.path.abspath(ckpt_dir)}")

    @torch.no_grad()
    def __call__(self, prompts: List[str]) -> Tuple[None, torch.Tensor, None]:
        hit = [self._cache.get(p) for p in prompts]
        if all(h is not None for h in hit):
            return None, torch.stack(hit).to(self.device), None
        ids, mask = _tokenize(self.tokenizer, prompts, True)
        enc = self.text_encoder(ids.to(self.device), attention_mask=mask.to(self.device))[0]
        enc = torch.nn.functional.normalize(enc, dim=-1).to(dtype=self.text_encoder.dtype, device=self.device)
        for p, e in zip(prompts, enc):
            self._cache.put(p, e.cpu())
        return None, enc, None


class AudioLDM2ProjectionModel(torch.nn.Module):
    """[UPSTREAM] diffusers AudioLDM2ProjectionModel.forward (text-to-audio variant, no learned position embedding):
    project each stream to the language-model width, wrap it in its learned SOS / EOS embeddings (mask extended with
    ones), concatenate along the sequence axis.  Parameter names follow the diffusers state dict."""

    def __init__(self, text_encoder_dim: int, text_encoder_1_dim: int, langauge_model_dim: int):
        super().__init__()
        self.projection = torch.nn.Linear(text_encoder_dim, langauge_model_dim)
        self.projection_1 = torch.nn.Linear(text_encoder_1_dim, langauge_model_dim)
        self.sos_embed = torch.nn.Parameter(torch.ones(langauge_model_dim))
        self.eos_embed = torch.nn.Parameter(torch.ones(langauge_model_dim))
        self.sos_embed_1 = torch.nn.Parameter(torch.ones(langauge_model_dim))
        self.eos_embed_1 = torch.nn.Parameter(torch.ones(langauge_model_dim))

    @staticmethod
    def _wrap(hs, mask, sos, eos):
        B = hs.shape[0]
        if mask is not None:
            one = mask.new_ones((B, 1))
            mask = torch.cat([one, mask, one], dim=-1)
        hs = torch.cat([sos.expand(B, 1, -1), hs, eos.expand(B, 1, -1)], dim=1)
        return hs, mask

    def forward(self, hidden_states, hidden_states_1, attention_mask, attention_mask_1):
        hs, m = self._wrap(self.projection(hidden_states), attention_mask, self.sos_embed, self.eos_embed)
        hs1, m1 = self._wrap(self.projection_1(hidden_states_1), attention_mask_1, self.sos_embed_1, self.eos_embed_1)
        return torch.cat([hs, hs1], dim=1), torch.cat([m, m1], dim=-1)


def generate_language_model(language_model, inputs_embeds: torch.Tensor, attention_mask: torch.Tensor,
                            max_new_tokens: Optional[int] = None) -> torch.Tensor:
    """[UPSTREAM] AudioLDM2Pipeline.generate_language_model: `max_new_tokens` (config value, 8) steps of GPT2Model on
    continuous inputs; each step appends the last position's hidden state to the input sequence and a one to the
    mask.  The pipeline carries a KV cache; recomputing the causal prefix gives the same values."""
    n = max_new_tokens if max_new_tokens is not None else getattr(language_model.config, "max_new_tokens", 8)
    for _ in range(n):
        out = language_model(inputs_embeds=inputs_embeds, attention_mask=attention_mask, use_cache=False,
                             return_dict=True).last_hidden_state
        inputs_embeds = torch.cat([inputs_embeds, out[:, -1:, :]], dim=1)
        attention_mask = torch.cat([attention_mask, attention_mask.new_ones((attention_mask.shape[0], 1))], dim=-1)
    return inputs_embeds[:, -n:, :]


class AudioLDM2TextStack:
    """AudioLDM2 conditioning (models.py:599-677)."""

    def __init__(self, ckpt_dir: str, device):
        from transformers import AutoTokenizer, ClapModel, GPT2Model, T5EncoderModel
        self.device = torch.device(device)
        sub = lambda n: os.path.join(ckpt_dir, n)
        self.tokenizer = AutoTokenizer.from_pretrained(sub("tokenizer"), local_files_only=True)
        self.tokenizer_2 = AutoTokenizer.from_pretrained(sub("tokenizer_2"), local_files_only=True)
        self.text_encoder = ClapModel.from_pretrained(sub("text_encoder"), local_files_only=True).to(self.device).eval()
        self.text_encoder_2 = T5EncoderModel.from_pretrained(sub("text_encoder_2"), local_files_only=True
                                                             ).to(self.device).eval()
        self.language_model = GPT2Model.from_pretrained(sub("language_model"), local_files_only=True
                                                        ).to(self.device).eval()
        pcfg = json.load(open(sub("projection_model/config.json")))
        self.projection_model = AudioLDM2ProjectionModel(pcfg["text_encoder_dim"], pcfg["text_encoder_1_dim"],
                                                         pcfg["langauge_model_dim"])
        wpath = _find(sub("projection_model"), "diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin")
        if wpath is None:
            raise FileNotFoundError(f"{sub('projection_model')}: no diffusion_pytorch_model.safetensors / .bin")
        missing, unexpected = self.projection_model.load_state_dict(_load_state(wpath), strict=False)
        if missing:
            raise KeyError(f"projection_model checkpoint lacks {missing}")
        self.projection_model.to(self.device).eval()
        self._cache = _DiskCache(f"audioldm2:{os.path.abspath(ckpt_dir)}")

    @torch.no_grad()
    def _encode(self, prompts: List[str]):
        dev = self.device
        # stream 0: CLAP pooled text features as a length-1 sequence that is always attended (models.py:629-637)
        ids, mask = _tokenize(self.tokenizer, prompts, True)
        clap = self.text_encoder.get_text_features(ids.to(dev), attention_mask=mask.to(dev))
        if not torch.is_tensor(clap):                       # newer transformers return an output object
            clap = clap.pooler_output if getattr(clap, "pooler_output", None) is not None else clap[0]
        clap = clap[:, None, :]
        clap_mask = mask.new_ones((len(prompts), 1)).to(dev)
        # stream 1: T5 encoder hidden states, padding to the longest prompt of the call (models.py:608,639-643)
        ids2, mask2 = _tokenize(self.tokenizer_2, prompts, False)
        mask2 = mask2.to(dev)
        t5 = self.text_encoder_2(ids2.to(dev), attention_mask=mask2)[0]
        proj, proj_mask = self.projection_model(clap, t5, clap_mask, mask2)
        gen = generate_language_model(self.language_model, proj, proj_mask, None)
        return (gen.to(dtype=self.language_model.dtype), t5.to(dtype=self.text_encoder_2.dtype), mask2)

    def __call__(self, prompts: List[str]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        # T5 pads to the longest prompt OF THE CALL and the padded positions take part in the GPT-2 prefix (masked), so
        # the cache key is the whole prompt list of a call
        key = "\x1f".join(prompts)
        hit = self._cache.get(key)
        if hit is not None:
            return tuple(t.to(self.device) for t in hit)
        out = self._encode(prompts)
        self._cache.put(key, tuple(t.cpu() for t in out))
        return out


class TangoTextEncoder:
    """TANGO conditioning (models.py:455-460 -> tango AudioDiffusion.encode_text [UPSTREAM]): T5 tokenizer with
    padding=True / truncation at model_max_length, frozen T5 encoder, boolean mask."""

    def __init__(self, ckpt_dir: str, device):
        from transformers import AutoTokenizer, T5Config, T5EncoderModel
        self.device = torch.device(device)
        main_cfg = {}
        p = os.path.join(ckpt_dir, "main_config.json")
        if os.path.exists(p):
            main_cfg = json.load(open(p))
        name = main_cfg.get("text_encoder_name", "google/flan-t5-large")
        tok_dir = os.path.join(ckpt_dir, "tokenizer")
        self.tokenizer = AutoTokenizer.from_pretrained(tok_dir if os.path.isdir(tok_dir) else name, local_files_only=True)
        enc_dir = os.path.join(ckpt_dir, "text_encoder")
        if os.path.isdir(enc_dir):
            self.text_encoder = T5EncoderModel.from_pretrained(enc_dir, local_files_only=True)
        else:
            # the snapshot the reference downloads keeps the encoder inside pytorch_model_main.bin (models.py:418-422)
            main = _find(ckpt_dir, "pytorch_model_main.bin", "pytorch_model_main.safetensors")
            if main is None:
                raise FileNotFoundError(f"{ckpt_dir}: neither text_encoder/ nor pytorch_model_main.bin")
            cfg = T5Config.from_pretrained(name, local_files_only=True)
            self.text_encoder = T5EncoderModel(cfg)
            sd = {k[len("text_encoder."):]: v for k, v in _load_state(main).items() if k.startswith("text_encoder.")}
            self.text_encoder.load_state_dict(sd)
        self.text_encoder.to(self.device).eval()
        self._cache = _DiskCache(f"tango:{os.path.abspath(ckpt_dir)}")

    @torch.no_grad()
    def __call__(self, prompts: List[str]) -> Tuple[torch.Tensor, None, torch.Tensor]:
        key = "\x1f".join(prompts)
        hit = self._cache.get(key)
        if hit is not None:
            return hit[0].to(self.device), None, hit[1].to(self.device)
        batch = self.tokenizer(prompts, max_length=self.tokenizer.model_max_length, padding=True, truncation=True,
                               return_tensors="pt")
        ids, mask = batch.input_ids.to(self.device), batch.attention_mask.to(self.device)
        hs = self.text_encoder(input_ids=ids, attention_mask=mask)[0]
        bmask = (mask == 1)
        self._cache.put(key, (hs.cpu(), bmask.cpu()))
        return hs, None, bmask


def has_text_checkpoint(ckpt_dir: Optional[str], family: str) -> bool:
    if not ckpt_dir or not os.path.isdir(ckpt_dir):
        return False
    if family == "tango":
        return (os.path.isdir(os.path.join(ckpt_dir, "text_encoder"))
                or _find(ckpt_dir, "pytorch_model_main.bin", "pytorch_model_main.safetensors") is not None)
    return os.path.isdir(os.path.join(ckpt_dir, "text_encoder")) and os.path.isdir(os.path.join(ckpt_dir, "tokenizer"))


def load_text_stack(ckpt_dir: str, family: str, device):
    return {"audioldm": ClapTextEncoder, "audioldm2": AudioLDM2TextStack, "tango": TangoTextEncoder}[family](ckpt_dir, device)
