"""Per-launch times of one MNIST-UNet forward at B=4096 (CUDA events around every launch; [prof] lines on stderr).
   python tools/layer_times.py [conv_stack 0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200 import _lib  # noqa: E402
from composable_diffusion_models_b200.models import UNet  # noqa: E402

stack = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 28
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
lib = _lib.lib()
lib.cdm_set_option(b"conv_stack", stack)
if len(sys.argv) > 4:
    lib.cdm_set_option(b"conv_pair", int(sys.argv[4]))
if len(sys.argv) > 5:
    lib.cdm_set_option(b"conv_pair64", int(sys.argv[5]))
if len(sys.argv) > 6:
    lib.cdm_set_option(b"stack_pair", int(sys.argv[6]))
m = UNet(precision="fp16").cuda().eval()
x = torch.randn(B, 1, S, S, device="cuda")
t = torch.full((B,), 0.5, device="cuda")
for _ in range(3):
    m(x, t)
torch.cuda.synchronize()
print(f"== conv_stack={stack} S={S} B={B}", file=sys.stderr, flush=True)
_lib.prof_enable(True)
m(x, t)
torch.cuda.synchronize()
lib.cdm_prof_dump()
_lib.prof_enable(False)
