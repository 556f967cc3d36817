"""Per-launch times of one GuidedUNet forward at B=2048, 32x32 (CUDA events around every launch; [prof] lines on stderr)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200 import _lib  # noqa: E402
from composable_diffusion_models_b200.models import GuidedUNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lib = _lib.lib()
g = GuidedUNet(precision="fp16").cuda().eval()
x = torch.randn(B, 3, 32, 32, device="cuda"); tt = torch.full((B,), 250.0, device="cuda")
d = torch.full((B,), 7, device="cuda"); c = torch.full((B,), 2, device="cuda")
for _ in range(3):
    g(x, tt, d, c)
torch.cuda.synchronize()
_lib.prof_enable(True)
g(x, tt, d, c)
torch.cuda.synchronize()
lib.cdm_prof_dump()
_lib.prof_enable(False)
