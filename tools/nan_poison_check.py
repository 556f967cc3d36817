"""Debug: run one test function after poisoning the CUDA caching allocator's free blocks with NaNs, so that any read of
memory the library never wrote shows up as a NaN / wrong result instead of depending on what ran before.
   python tools/nan_poison_check.py tests.test_gpu_ito test_unet_jvp_matches_autograd_vjp 3 64 fp32 1e-5 1e-4"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

mod, fn = sys.argv[1], sys.argv[2]


def conv(a):
    try:
        return int(a)
    except ValueError:
        try:
            return float(a)
        except ValueError:
            return a


args = [conv(a) for a in sys.argv[3:]]
for fill in (float("nan"), 1e30, -7.0):
    blocks = [torch.full((1 << 28,), fill, device="cuda") for _ in range(8)]   # 8 GiB of poison
    small = [torch.full((n,), fill, device="cuda") for n in (1 << 10, 1 << 14, 1 << 18, 1 << 22) for _ in range(16)]
    del blocks, small
    getattr(importlib.import_module(mod), fn)(*args)
    print("ok with poison", fill, flush=True)
