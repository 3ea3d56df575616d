"""Shim: `from models import load_model, PipelineWrapper` (what code/main_run*.py, main_pc_*.py do) resolves here."""
from audioeditingcode_b200.models import *  # noqa: F401,F403
from audioeditingcode_b200.models import PipelineWrapper, AudioLDMWrapper, AudioLDM2Wrapper, TangoWrapper, load_model  # noqa: F401
