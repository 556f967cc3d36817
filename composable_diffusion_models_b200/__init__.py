"""composable_diffusion_models_b200 -- B200-native (sm_100a) composed-score diffusion sampling.

Drop-in for the sampler hot path of mo-rsa24/composable_diffusion_models: the
modules here keep the reference's names, call signatures and checkpoint
state_dict keys; the work runs in hand-written CUDA behind the C ABI of
``libcdm_b200.so`` (include/cdm_b200.h).  No CPU fallback exists.
"""
__version__ = "0.1.0"

from . import schedule  # noqa: F401
