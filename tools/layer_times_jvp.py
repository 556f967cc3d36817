"""Per-launch times of one UNet forward_jvp (primal + tangent, fp16) at B=1024, 64x64 ([prof] lines on stderr)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200 import _lib  # noqa: E402
from composable_diffusion_models_b200.models import UNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = _lib.lib()
m = UNet(in_channels=3, num_classes=3, precision="fp16").cuda().eval()
x = torch.randn(B, 3, 64, 64, device="cuda"); v = torch.randn_like(x)
t = torch.full((B,), 0.5, device="cuda"); y = torch.full((B,), 1, device="cuda")
for _ in range(2):
    m.forward_jvp(x, t, y, v)
torch.cuda.synchronize()
_lib.prof_enable(True)
m.forward_jvp(x, t, y, v)
torch.cuda.synchronize()
lib.cdm_prof_dump()
_lib.prof_enable(False)
