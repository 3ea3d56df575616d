"""Shim for `from ddm_inversion.ddim_inversion import ddim_inversion, text2image_ldm_stable`."""
from audioeditingcode_b200.ddm_inversion.ddim_inversion import ddim_inversion, text2image_ldm_stable, next_step, get_noise_pred  # noqa: F401
