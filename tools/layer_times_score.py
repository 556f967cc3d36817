"""Per-launch times of one BatchNorm score-UNet forward (fp32 path) at B=1024, 32x32 ([prof] lines on stderr)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200 import _lib  # noqa: E402
from composable_diffusion_models_b200.models import ColoredMNISTScoreModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = _lib.lib()
m = ColoredMNISTScoreModel().cuda().eval()
x = torch.randn(B, 3, 32, 32, device="cuda"); t = torch.full((B,), 500.0, device="cuda")
for _ in range(2):
    m(x, t)
torch.cuda.synchronize()
_lib.prof_enable(True)
m(x, t)
torch.cuda.synchronize()
lib.cdm_prof_dump()
_lib.prof_enable(False)
