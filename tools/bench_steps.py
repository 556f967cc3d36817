"""Microbenchmark of the fused combine+update step kernels (HBM roofline): GB/s of algorithmic bytes per mode at the
BASELINE.json shapes, L2 flushed between timed launches (a 512 MB memset), CUDA events on the launching stream."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from composable_diffusion_models_b200 import steps as S

dev = "cuda"
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


REPS = int(os.environ.get("STEPS_REPS", "20"))      # STEPS_REPS=1: one warm + one timed launch per case (ncu target)


def timeit(fn, nbytes, reps=REPS):
    for _ in range(3 if reps > 1 else 1):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    return ms, nbytes / ms / 1e6


rows = []
def case(name, fn, nbytes):
    ms, gbs = timeit(fn, nbytes)
    rows.append(dict(kernel=name, ms=round(ms, 4), algorithmic_MB=round(nbytes / 1e6, 1), GBps=round(gbs, 1), frac_of_measured_hbm=round(gbs / peak, 3)))
    print(rows[-1], flush=True)

# C2: MNIST K=2 SDE, B=4096 (and 8x that), injected noise vs in-kernel Philox
for B in (4096, 32768):
    x = torch.randn(B, 1, 28, 28, device=dev); e1 = torch.randn_like(x); e2 = torch.randn_like(x); z = torch.randn_like(x)
    D = 784 * 4
    case(f"sde K=2 B={B} 1x28x28 injected z", lambda: S.step_sde(x, [e1, e2], [1.0, 1.0], -5.0, 3.0, 1e-3, 0.1, z=z, out=x), B * D * 5)
    case(f"sde K=2 B={B} 1x28x28 in-kernel rng", lambda: S.step_sde(x, [e1, e2], [1.0, 1.0], -5.0, 3.0, 1e-3, 0.1, rng=(1, 2), out=x), B * D * 4)
# C3: shapes DDIM, B=8192, 3x64x64, shape expert 1 channel, gray out
B = 8192
x = torch.randn(B, 3, 64, 64, device=dev); es = torch.randn(B, 1, 64, 64, device=dev); ec = torch.randn_like(x); gray = torch.empty(B, 1, 64, 64, device=dev)
case("ddim K=2 B=8192 3x64x64 (+gray)", lambda: S.step_ddim(x, [es, ec], [1.0, 1.0], 2.0, 0.5, 0.8, 0.6, 0.7, out=x, gray_out=gray), B * 4096 * 4 * (3 + 3 + 1 + 3 + 1))
# C4: SuperDiff K=4 log-q, B=1024 3x64x64 and B=8192 3x32x32
for B, Sz in ((1024, 64), (8192, 32)):
    x = torch.randn(B, 3, Sz, Sz, device=dev); ns = [torch.randn_like(x) for _ in range(4)]; z = torch.randn_like(x); lq = torch.zeros(B, 4, device=dev)
    case(f"ddpm_logq K=4 B={B} 3x{Sz}x{Sz}", lambda: S.step_ddpm_logq(x, ns, lq, "OR", 1.0, 0.0, 0.9, 0.01, 0.99, 0.05, 1e-3, z=z, out=x), B * 3 * Sz * Sz * 4 * 7)
    d1 = torch.randn(B, device=dev); d2 = torch.randn(B, device=dev); e1 = torch.randn(B, 1, Sz, Sz, device=dev)
    case(f"ode_kappa B={B} 3x{Sz}x{Sz}", lambda: S.step_ode_kappa(x, e1, ns[0], d1, d2, 0.9, -5.0, 2.0, 1e-3, div1_scale=3.0, out=x), B * Sz * Sz * 4 * (3 + 3 + 1 + 3))
# C5: CFG x0 form, B=2048 3x32x32
B = 2048
x = torch.randn(B, 3, 32, 32, device=dev); ps = [torch.randn_like(x) for _ in range(3)]
case("cfg K=3 B=2048 3x32x32", lambda: S.step_cfg(x, ps, [1.0, 7.5, 7.5], 1.0, 0, 0, 0.9, 0.4, out=x), B * 3072 * 4 * 4)
if REPS > 1:
    json.dump(rows, open("gpurun_out/bench_steps.json", "w"), indent=1)
