"""deepinpainting_b200 -- B200 (sm_100a) implementation of DeepInPainting's IPSR / CSA patch-shift
attention layer behind the reference's own Python API.

    from deepinpainting_b200.models import IPSR_model, IPSRFunction, InnerCos, InnerCos2
    from deepinpainting_b200.util import util, NonparametricShift, MaxCoord

Host code is Python/PyTorch plumbing over the C ABI of ``lib/libipsr_sm100.so``
(``include/ipsr_sm100.h``); there is no CPU fallback.
"""
from . import _lib, build, shift_ops        # noqa: F401

__all__ = ["_lib", "build", "shift_ops", "models", "util"]
