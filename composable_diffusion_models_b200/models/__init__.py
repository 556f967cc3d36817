"""Expert denoisers with the reference's constructor/forward signatures, executed by libcdm_b200."""
from .unet_small import UNet, forward_grouped  # noqa: F401
from .mlp_2d import MLP  # noqa: F401
from .score_model import ColoredMNISTScoreModel, ScoreModel  # noqa: F401
from .guided_unet import GuidedUNet  # noqa: F401
from .beta_vae import BetaVAE, quantize_u8  # noqa: F401
from .simple_unet import SimpleUnet  # noqa: F401
