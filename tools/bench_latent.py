"""C1 (BASELINE.json configs[0]): 2-D latent MLP experts, K=2 SDE composition, whole 1000-step chain in one persistent
launch.  python tools/bench_latent.py [log2_B]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from composable_diffusion_models_b200.compose_scores import sample_composed_latent_sde  # noqa: E402
from composable_diffusion_models_b200.models import MLP  # noqa: E402

lb = int(sys.argv[1]) if len(sys.argv) > 1 else 17
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
B = 1 << lb
torch.manual_seed(0)
experts = [MLP().cuda().eval() for _ in range(2)]
for _ in range(2):
    sample_composed_latent_sde(experts, [1.0, 1.0], 4096, 50, noise="kernel", seed=1, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
x = sample_composed_latent_sde(experts, [1.0, 1.0], B, 1000, noise="kernel", seed=2, precision=prec)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
gflop = B * 2 * 1000 * 0.265e-3
print({"config": f"C1 latent MLP K=2 SDE, 1000 steps, persistent kernel ({'fp16 tcgen05' if prec == 'fp16' else 'fp32 CUDA cores'})", "batch": B, "ms": round(ms, 1),
       "samples_per_s": round(B / ms * 1e3), "tflops": round(gflop / ms, 1), "finite": bool(torch.isfinite(x).all())})
