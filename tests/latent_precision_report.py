import sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import rel_l2
from oracle import experts as E, samplers as OS
from composable_diffusion_models_b200.compose_scores import sample_composed_latent_sde
from composable_diffusion_models_b200.models import MLP
sds = [E.synth_state_dict(E.mlp_2d_spec(), s) for s in (41, 42)]
ms = []
for sd in sds:
    m = MLP(); m.load_state_dict(sd, strict=True); ms.append(m.cuda())
for n_steps in (60, 1000):
    g = torch.Generator().manual_seed(4)
    B = 256
    x0 = torch.randn(B, 2, generator=g); noise = torch.randn(n_steps, B, 2, generator=g)
    want = OS.sample_sde([lambda x, t, sd=sd: E.mlp_2d_forward(sd, t, x) for sd in sds], [1.0, 0.5], x0, noise, n_steps, 1.0)
    for prec in ("fp32", "fp16"):
        got = sample_composed_latent_sde(ms, [1.0, 0.5], B, n_steps, 1.0, device="cuda", x_init=x0, noise=noise, precision=prec)
        print("latent chain", n_steps, prec, rel_l2(got.cpu(), want))
