"""Shim for `from ddm_inversion.inversion_utils import inversion_forward_process, inversion_reverse_process`."""
from audioeditingcode_b200.ddm_inversion.inversion_utils import inversion_forward_process, inversion_reverse_process  # noqa: F401
