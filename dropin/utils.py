"""Shim for `from utils import load_audio, set_reproducability, get_text_embeddings, get_spec, ...`."""
from audioeditingcode_b200.utils import *  # noqa: F401,F403
from audioeditingcode_b200.utils import load_audio, get_spec, set_reproducability, get_height_of_spectrogram, get_text_embeddings  # noqa: F401
