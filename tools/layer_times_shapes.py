"""Per-launch times of one shapes-UNet forward (3x64x64, conditional) at B=4096 ([prof] lines on stderr).
   python tools/layer_times_shapes.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200 import _lib  # noqa: E402
from composable_diffusion_models_b200.models import UNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.lib()
m = UNet(in_channels=3, num_classes=3, precision="fp16").cuda().eval()
x = torch.randn(B, 3, 64, 64, device="cuda")
t = torch.full((B,), 0.5, device="cuda")
y = torch.full((B,), 1, device="cuda")
for _ in range(3):
    m(x, t, y)
torch.cuda.synchronize()
_lib.prof_enable(True)
m(x, t, y)
torch.cuda.synchronize()
lib.cdm_prof_dump()
_lib.prof_enable(False)
