"""Shared builders for the full-size parity tests / tools (test infrastructure: may import oracle/).

The benchmarked geometries (BASELINE configs[1] AudioLDM2-large and configs[2] TANGO-full, 10 s clip -> latent
[*, 8, 256, 16]) with seeded synthetic weights and seeded text conditioning, and the fp32 CPU oracle U-Net bound to
the same weights / text (oracle/unet_torch.py; reference code/models.py:160-393, :691-899)."""
import torch

from oracle import unet_torch as U
from audioeditingcode_b200 import unet_config as C

MODEL_IDS = {"audioldm2-large": "cvssp/audioldm2-large", "tango": "declare-lab/tango-full-ft-audiocaps",
             "audioldm2": "cvssp/audioldm2", "audioldm-s": "cvssp/audioldm-s-full-v2"}


def text_rows(cfg, n_rows, lens=(8, 16), seed=4, uncond_first=True):
    """Seeded text conditioning rows in the engine's convention: (streams, masks, class_labels).  Row 0 is the
    unconditional row when uncond_first (stream 1 / TANGO stream: a single valid token, the rest masked)."""
    gen = torch.Generator().manual_seed(seed)
    if cfg.class_embed_dim is not None:
        y = torch.nn.functional.normalize(torch.randn(n_rows, cfg.class_embed_dim, generator=gen), dim=-1)
        return [], [], y
    dims = {s[1]: s[0] for s in cfg.transformer_specs if s is not None}
    streams, masks = [], []
    for i in range(cfg.n_streams):
        masked = (i == cfg.n_streams - 1)          # the T5 stream carries the mask (models.py:706-710)
        L = lens[1] if masked else lens[0]
        streams.append(torch.randn(n_rows, L, dims[i], generator=gen))
        m = torch.ones(n_rows, L)
        if masked and uncond_first:
            m[0, 1:] = 0
        masks.append(m if masked else None)
    return streams, masks, None


def oracle_eval(cfg, w, x, t, streams, masks, cl, rows, probe=None):
    """fp32 CPU oracle evaluation of samples x [B,...] whose text rows are `rows` (list of row indices)."""
    idx = torch.as_tensor(rows, dtype=torch.long)
    kw = {}
    if cl is not None:
        kw["class_labels"] = cl[idx]
    if streams:
        kw["streams"] = [s[idx] for s in streams]
        kw["stream_masks"] = [None if m is None else m[idx] for m in masks]
    tt = torch.as_tensor(t, dtype=torch.int64).reshape(-1).expand(x.shape[0]) if not torch.is_tensor(t) or t.numel() == 1 \
        else t
    with torch.no_grad():
        return U.unet_forward(cfg, w, x, tt, probe=probe, **kw)[0]


class RowText:
    """encode_text stub for a wrapper: maps prompt strings to fixed rows of seeded conditioning, returned in the
    REFERENCE's (encoder_hidden_states, class_labels, encoder_attention_mask) convention of each family
    (models.py:455-460, 511-537, 599-677)."""

    def __init__(self, cfg, family, streams, masks, cl, prompt_rows, device):
        self.cfg, self.family, self.streams, self.masks, self.cl = cfg, family, streams, masks, cl
        self.prompt_rows, self.device = prompt_rows, device

    def __call__(self, prompts, **kw):
        idx = torch.as_tensor([self.prompt_rows[p] for p in prompts], dtype=torch.long)
        dev = self.device
        if self.family == "audioldm":
            return None, self.cl[idx].to(dev), None
        if self.family == "audioldm2":
            return self.streams[0][idx].to(dev), self.streams[1][idx].to(dev), self.masks[1][idx].to(dev)
        return self.streams[0][idx].to(dev), None, self.masks[0][idx].to(dev).bool()


def family_of(name):
    return "tango" if "tango" in name else ("audioldm2" if "audioldm2" in name else "audioldm")
