"""Measured rel-L2 error of the fp16 tensor-core path against the fp32 oracle: single forwards (each conv variant)
and composed SDE chains -- next to the error of the reference's OWN default GPU path (torch conv2d with cuDNN TF32
enabled, which is what the reference runs on a GPU), measured the same way.  Test-support script (it imports the
oracle, so it lives under tests/).  Run on the GPU box:  python tests/precision_report.py [out.json]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from composable_diffusion_models_b200 import _lib  # noqa: E402
from composable_diffusion_models_b200.compose_scores import sample_composed_sde  # noqa: E402
from composable_diffusion_models_b200.models import UNet  # noqa: E402
from oracle import experts as E  # noqa: E402
from oracle import samplers as OS  # noqa: E402


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def unet(kw, seed, precision):
    m = UNet(**kw, precision=precision)
    sd = E.synth_state_dict(E.unet_small_spec(kw.get("in_channels", 1), num_classes=kw.get("num_classes")), seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


rows = []
lib = _lib.lib()
for cin, S in ((1, 28), (3, 64), (1, 64)):
    nc = 3 if cin == 3 else None
    m, sd = unet(dict(in_channels=cin, num_classes=nc), 321, "fp16")
    g = torch.Generator().manual_seed(5)
    B = 5
    x = torch.randn(B, cin, S, S, generator=g)
    t = torch.rand(B, generator=g) * 0.9 + 0.05
    y = torch.randint(0, 3, (B,), generator=g) if nc else None
    want = E.unet_small_forward(sd, x, t, y)
    for name, halo, fuse in (("box", 0, 0), ("halo", 1, 0), ("halo+gn", 1, 1)):
        lib.cdm_set_option(b"conv_halo", halo)
        lib.cdm_set_option(b"fuse_gn", fuse)
        got = m(x.cuda(), t.cuda(), y.cuda() if nc else None).cpu()
        rows.append(dict(kind="forward", cin=cin, S=S, variant=name, rel_l2=rel_l2(got, want)))
        print(rows[-1], flush=True)
    lib.cdm_set_option(b"conv_halo", -1)
    lib.cdm_set_option(b"fuse_gn", -1)

for n_steps, B in ((40, 3), (200, 3), (1000, 2)):
    seeds = (301, 302)
    sds = [E.synth_state_dict(E.unet_small_spec(1), s) for s in seeds]
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(B, 1, 28, 28, generator=g)
    noise = torch.randn(n_steps, B, 1, 28, 28, generator=g)
    want = OS.sample_sde([lambda x, t, sd=sd: E.unet_small_forward(sd, x, t) for sd in sds], [1.0, 1.0], x0, noise, n_steps, 1.0)
    for prec in ("fp32", "fp16"):
        experts = [unet(dict(in_channels=1), s, prec)[0] for s in seeds]
        got = sample_composed_sde(experts, [1.0, 1.0], B, (1, 28, 28), n_steps, 1.0, device="cuda", x_init=x0, noise=noise)
        rows.append(dict(kind="sde_chain", n_steps=n_steps, B=B, precision=prec, rel_l2=rel_l2(got.cpu(), want)))
        print(rows[-1], flush=True)
    # the same chain with w1 + w2 = 1 (a normalised mixture, as the shapes samplers use)
    want = OS.sample_sde([lambda x, t, sd=sd: E.unet_small_forward(sd, x, t) for sd in sds], [0.5, 0.5], x0, noise, n_steps, 1.0)
    experts = [unet(dict(in_channels=1), s, "fp16")[0] for s in seeds]
    got = sample_composed_sde(experts, [0.5, 0.5], B, (1, 28, 28), n_steps, 1.0, device="cuda", x_init=x0, noise=noise)
    rows.append(dict(kind="sde_chain_w0.5", n_steps=n_steps, B=B, precision="fp16", rel_l2=rel_l2(got.cpu(), want)))
    print(rows[-1], flush=True)

# ---- the reference's default GPU arithmetic: the same torch ops on CUDA tensors, conv2d in TF32 (torch default) ----
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = False   # torch default: Linear stays fp32


def cuda_sd(sd):
    return {k: v.cuda() for k, v in sd.items()}


for cin, S in ((1, 28), (3, 64)):
    nc = 3 if cin == 3 else None
    sd = E.synth_state_dict(E.unet_small_spec(cin, num_classes=nc), 321)
    g = torch.Generator().manual_seed(5)
    B = 5
    x = torch.randn(B, cin, S, S, generator=g)
    t = torch.rand(B, generator=g) * 0.9 + 0.05
    y = torch.randint(0, 3, (B,), generator=g) if nc else None
    want = E.unet_small_forward(sd, x, t, y)
    got = E.unet_small_forward(cuda_sd(sd), x.cuda(), t.cuda(), y.cuda() if nc else None).cpu()
    rows.append(dict(kind="forward_torch_cuda_tf32", cin=cin, S=S, rel_l2=rel_l2(got, want)))
    print(rows[-1], flush=True)

for n_steps, B in ((40, 3), (200, 3)):
    seeds = (301, 302)
    sds = [E.synth_state_dict(E.unet_small_spec(1), s) for s in seeds]
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(B, 1, 28, 28, generator=g)
    noise = torch.randn(n_steps, B, 1, 28, 28, generator=g)
    want = OS.sample_sde([lambda x, t, sd=sd: E.unet_small_forward(sd, x, t) for sd in sds], [1.0, 1.0], x0, noise, n_steps, 1.0)
    csds = [cuda_sd(sd) for sd in sds]
    got = OS.sample_sde([lambda x, t, sd=sd: E.unet_small_forward(sd, x.cuda(), t.cuda()).cpu() for sd in csds], [1.0, 1.0], x0, noise,
                        n_steps, 1.0)
    rows.append(dict(kind="sde_chain_torch_cuda_tf32", n_steps=n_steps, B=B, rel_l2=rel_l2(got, want)))
    print(rows[-1], flush=True)

if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
