"""Per-step time of the other BASELINE.json configurations (C3 shapes DDIM, C4 SuperDiff K=4, C5 GuidedUNet CFG) with
synthetic weights -- orientation numbers for DESIGN.md, not the headline metric."""
import json, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from composable_diffusion_models_b200 import steps as S, schedule
from composable_diffusion_models_b200.models import UNet, ColoredMNISTScoreModel, GuidedUNet
from composable_diffusion_models_b200.compose_images_ddim import ddim_tables

dev = "cuda"
rows = []

def timed(fn, steps=5, warm=2):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): fn(warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

# C3: shapes 64x64, two-expert DDIM, B = 8192, 50 steps
B = int(os.environ.get("C3_B", "8192"))
torch.manual_seed(0)
ms_ = UNet(in_channels=1, num_classes=3).to(dev).eval(); mc_ = UNet(in_channels=3, num_classes=3).to(dev).eval()
x = torch.randn(B, 3, 64, 64, device=dev); xg = S.grayscale(x)
sl = torch.full((B,), 2, device=dev); cl = torch.full((B,), 1, device=dev)
ts, al, sg = [v.tolist() for v in ddim_tables(50)]
tv = torch.empty(B, device=dev)
def c3(i):
    global x
    i = i % 50
    tv.fill_(ts[i])
    es = ms_(xg, tv, sl); ec = mc_(x, tv, cl)
    S.step_ddim(x, [es, ec], [1.0, 1.0], 2.0, al[i], sg[i], al[i + 1], sg[i + 1], out=x, gray_out=xg)
ms = timed(c3)
gflop = 4.166 + 4.177
rows.append(dict(config="C3 shapes 64x64 DDIM K=2 (fp16 tcgen05)", batch=B, ms_per_step=round(ms, 2), samples_per_s=round(B / (50 * ms * 1e-3), 1),
                 tflops=round(B * gflop / ms, 1)))
print(rows[-1], flush=True)
del ms_, mc_, x, xg
torch.cuda.empty_cache()

# C4 (as BASELINE.json words it): shapes 64x64 Ito kappa-ODE, shape + colour experts, Hutchinson divergence by forward-mode
# JVP (primal + tangent through both UNets every step), B = 1024
from composable_diffusion_models_b200 import compose_images_ito as ITO
for prec, Bi in (("fp16", 1024), ("fp32", 64)):
    torch.manual_seed(0)
    ms_ = UNet(in_channels=1, num_classes=3, precision=prec).to(dev).eval()
    mc_ = UNet(in_channels=3, num_classes=3, precision=prec).to(dev).eval()
    xi = torch.randn(Bi, 3, 64, 64, device=dev)
    sli = torch.full((Bi,), 2, device=dev); cli = torch.full((Bi,), 1, device=dev)
    tvi = torch.empty(Bi, device=dev)
    def c4ito(i):
        global xi
        t_val = 1.0 - (i % 1000) * 1e-3
        tvi.fill_(t_val)
        xg = S.grayscale(xi)
        es, ds = ms_.forward_jvp(xg, tvi, sli, torch.randn_like(xg))
        ec, dc = mc_.forward_jvp(xi, tvi, cli, torch.randn_like(xi))
        S.step_ode_kappa(xi, es, ec, ds, dc, float(schedule.sigma(torch.tensor(t_val))), float(schedule.dlog_alphadt(torch.tensor(t_val))),
                         0.5 * float(schedule.beta(torch.tensor(t_val))), 1e-3, mode=0, div1_scale=3.0, out=xi)
    ms = timed(c4ito, steps=3, warm=2)
    rows.append(dict(config=f"C4 shapes 64x64 Ito kappa-ODE K=2, JVP divergence ({'fp16 tcgen05' if prec == 'fp16' else 'fp32 CUDA-core path'})",
                     batch=Bi, ms_per_step=round(ms, 2), samples_per_s=round(Bi / (1000 * ms * 1e-3), 3),
                     tflops=round(Bi * 2 * (4.166 + 4.177) / ms, 1)))
    print(rows[-1], flush=True)
    del ms_, mc_, xi
    torch.cuda.empty_cache()

# C4': SuperDiff, K = 4 BatchNorm score UNets (fp32 path), 3x32x32, B = 1024, T = 1000
B = 1024
experts = [ColoredMNISTScoreModel().to(dev).eval() for _ in range(4)]
x = torch.randn(B, 3, 32, 32, device=dev); lq = torch.zeros(B, 4, device=dev); tf = torch.empty(B, device=dev)
sde = schedule.VPSDE()
tb = sde.host_tables()
def c4(i):
    t_idx = 999 - i
    tf.fill_(float(t_idx))
    preds = [m(x, tf) for m in experts]
    S.step_ddpm_logq(x, preds, lq, "OR", 1.0, 0.0, float(tb["sqrt_one_minus_alphas_cumprod"][t_idx]), float(tb["betas"][t_idx]),
                     float(tb["alphas"][t_idx]) ** 0.5, float(tb["posterior_variance"][t_idx]) ** 0.5, 1e-3, rng=(0, i), out=x)
ms = timed(c4)
rows.append(dict(config="C4 SuperDiff K=4 score UNets 3x32x32 (fp32 CUDA-core path)", batch=B, ms_per_step=round(ms, 2),
                 samples_per_s=round(B / (1000 * ms * 1e-3), 2), tflops=round(B * 4 * 0.663 / ms, 1)))
print(rows[-1], flush=True)
del experts
torch.cuda.empty_cache()

# C5: GuidedUNet CFG (3 conditioned forwards per step), 3x32x32, B = 2048 per GPU, 500 steps
B = 2048
c5_prec = os.environ.get("C5_PRECISION", "fp16")
g = GuidedUNet(precision=c5_prec).to(dev).eval()
x = torch.randn(B, 3, 32, 32, device=dev); tt = torch.empty(B, device=dev)
d = torch.full((B,), 7, device=dev); c = torch.full((B,), 2, device=dev); nd = torch.full((B,), 10, device=dev); nc = torch.full((B,), 3, device=dev)
def c5(i):
    tt.fill_(float(499 - i))
    pu = g(x, tt, nd, nc); ps = g(x, tt, d, nc); pc = g(x, tt, nd, c)
    S.step_cfg(x, [pu, ps, pc], [1.0, 7.5, 7.5], 1.0, 0, 0, 0.9, 0.4, out=x)
ms = timed(c5, steps=3, warm=2)
rows.append(dict(config=f"C5 GuidedUNet CFG 3 fwd/step 3x32x32 ({'fp16 tcgen05' if c5_prec == 'fp16' else 'fp32 CUDA-core path'})", batch=B, ms_per_step=round(ms, 2),
                 samples_per_s=round(B / (500 * ms * 1e-3), 2), tflops=round(B * 3 * 2.482 / ms, 1)))
print(rows[-1], flush=True)
del g
torch.cuda.empty_cache()

# section 8(f) row 4: the 62 M-parameter SimpleUnet (fp32 CUDA-core path), one forward at 64x64
from composable_diffusion_models_b200.models import SimpleUnet
B = 256
su = SimpleUnet(3).to(dev).eval()
x = torch.randn(B, 3, 64, 64, device=dev); tt = torch.full((B,), 250.0, device=dev); yy = torch.full((B,), 1, device=dev)
ms = timed(lambda i: su(x, tt, yy), steps=3, warm=2)
rows.append(dict(config="SimpleUnet 62M params, one forward 3x64x64 (fp32 CUDA-core path)", batch=B, ms_per_step=round(ms, 2),
                 samples_per_s=round(B / (ms * 1e-3), 1), tflops=round(B * 11.46 / ms, 1)))
print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/bench_configs.json", "w"), indent=1)
