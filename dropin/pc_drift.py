"""Shim for `from pc_drift import forward_directional, get_eigenvectors, apply_drift, PromptEmbeddings, PCStreamChoice`."""
from audioeditingcode_b200.pc_drift import *  # noqa: F401,F403
from audioeditingcode_b200.pc_drift import PromptEmbeddings, PCStreamChoice, expand_for_evs, forward_directional, get_eigenvectors, apply_drift  # noqa: F401
