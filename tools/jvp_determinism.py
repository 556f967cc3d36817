"""Debug: repeat UNet forward_jvp on fixed inputs and report run-to-run deviations (float atomics give ~1e-7; anything larger
is a race).   python tools/jvp_determinism.py [precision] [S] [cin] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200.models import UNet  # noqa: E402
from oracle import experts as E  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cin = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 3
nc = 3
sd = E.synth_state_dict(E.unet_small_spec(cin, num_classes=nc), 900 + cin)
m = UNet(in_channels=cin, num_classes=nc, precision=prec)
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
g = torch.Generator().manual_seed(seed)
B = 3
x = torch.randn(B, cin, S, S, generator=g).cuda()
v = torch.randn(B, cin, S, S, generator=g).cuda()
t = (torch.rand(B, generator=g) * 0.9 + 0.05).cuda()
y = torch.randint(0, 3, (B,), generator=g).cuda()
eps0, div0 = m.forward_jvp(x, t, y, v)
fwd0 = m(x, t, y)
worst_e = worst_d = worst_f = 0.0
for i in range(reps):
    junk = torch.full((1 << 24,), float(i), device="cuda")
    del junk
    eps, div = m.forward_jvp(x, t, y, v)
    fwd = m(x, t, y)
    de = float((eps - eps0).abs().max()); dd = float((div - div0).abs().max() / div0.abs().max()); df = float((fwd - fwd0).abs().max())
    if (dd > 1e-5 or de > 1e-4 or df > 1e-4) and worst_d < 1e-5:
        print(f"rep {i}: eps dev {de:.3e}  div rel dev {dd:.3e}  fwd dev {df:.3e}  div={div.tolist()}")
    worst_e, worst_d, worst_f = max(worst_e, de), max(worst_d, dd), max(worst_f, df)
print(f"{prec} S={S} cin={cin} seed={seed}: worst eps dev {worst_e:.3e}, div rel dev {worst_d:.3e}, fwd dev {worst_f:.3e}")
