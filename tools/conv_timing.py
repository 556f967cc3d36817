"""Debug: per-role mbarrier wait cycles of every halo-conv launch of one UNet forward (B=4096, 28x28)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from composable_diffusion_models_b200 import _lib
from composable_diffusion_models_b200.models import UNet
fuse = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib = _lib.lib()
m = UNet(precision="fp16").cuda().eval()
x = torch.randn(4096, 1, 28, 28, device="cuda"); t = torch.full((4096,), 0.5, device="cuda")
lib.cdm_set_option(b"fuse_gn", fuse)
if len(sys.argv) > 2:
    lib.cdm_set_option(b"conv_stack", int(sys.argv[2]))
for _ in range(2): m(x, t)
torch.cuda.synchronize()
lib.cdm_set_option(b"conv_timing", 1)
m(x, t)
torch.cuda.synchronize()
