"""A/B of the init conv kernels inside whole UNet forwards (set CDM_INIT_CONV_TC = 0 | 1 | 2 in the environment):
rel-L2 of the fp16 forward against the fp32-mode forward of the same weights, and the init_conv launch time.
   python tools/init_conv_ab.py [cin] [S] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200 import _lib  # noqa: E402
from composable_diffusion_models_b200.models import UNet  # noqa: E402
from oracle import experts as E  # noqa: E402  (tools only: synthetic weights)

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 3
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
kw = dict(in_channels=cin, num_classes=3) if cin > 1 else dict(in_channels=1)
sd = E.synth_state_dict(E.unet_small_spec(cin, num_classes=kw.get("num_classes")), 7)
ms = {}
for prec in ("fp16", "fp32"):
    m = UNet(**kw, precision=prec)
    m.load_state_dict(sd, strict=True)
    ms[prec] = m.cuda().eval()
g = torch.Generator().manual_seed(3)
x = torch.randn(B, cin, S, S, generator=g).cuda()
t = torch.full((B,), 0.4, device="cuda")
args = (x, t, torch.full((B,), 1, device="cuda")) if cin > 1 else (x, t)
nb = min(B, 64)
small = tuple(a[:nb].contiguous() for a in args)
ref = ms["fp32"](*small)
out = ms["fp16"](*args)
out2 = ms["fp16"](*args)
rel = ((out[:nb] - ref).norm() / ref.norm()).item()
print(f"CDM_INIT_CONV_TC={os.environ.get('CDM_INIT_CONV_TC', '(default)')} cin={cin} S={S} B={B}: rel-L2 fp16 vs fp32 mode {rel:.3e}; "
      f"repeat bit-identical {bool((out == out2).all())}; finite {bool(torch.isfinite(out).all())}")
torch.cuda.synchronize()
_lib.prof_enable(True)
ms["fp16"](*args)
torch.cuda.synchronize()
_lib.lib().cdm_prof_dump()
_lib.prof_enable(False)
