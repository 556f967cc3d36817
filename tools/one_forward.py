"""Two MNIST-UNet forwards at B=4096 on the fp16 tcgen05 path (ncu target: the second forward is the warm one).
   ncu --set full --clock-control none --import-source on -k regex:conv_ -c 24 -o gpurun_out/r02_conv python tools/one_forward.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200.models import UNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 28
torch.manual_seed(1234)
m = UNet(precision="fp16").cuda().eval()
x = torch.randn(B, 1, S, S, device="cuda")
t = torch.full((B,), 0.5, device="cuda")
for _ in range(2):
    m(x, t)
torch.cuda.synchronize()
print("done")
