"""Plumbing shared by the whole-chain entry points (cdm_*_sample_*): one host call enqueues every sampler step, so there is
no Python between kernels.  The shims fall back to their per-step loops for experts that are not native modules."""
import ctypes as C

import torch

from . import _lib
from .models import _native


def native_all(models, cls, x=None):
    """Every model is exactly `cls` (a native module), in eval mode, with one precision."""
    if not models or not all(type(m) is cls and not m.training for m in models):
        return False
    p0 = getattr(models[0], "precision", None)
    if any(getattr(m, "precision", None) != p0 for m in models):
        return False
    return x is None or x.is_cuda


def handle_array(models, device):
    arr = (C.c_void_p * len(models))(*[m._native_handle(device).value for m in models])
    return arr, C.cast(arr, C.POINTER(C.c_void_p))


def label_arrays(labels, B, device):
    """labels: per expert None / int / tensor -> (host array of device pointers or None, keep-alive list, all uniform?)."""
    if labels is None or all(lab is None for lab in labels):
        return None, [], 1
    keep, uniform = [], True
    arr = (C.c_void_p * len(labels))()
    for i, lab in enumerate(labels):
        if lab is None:
            arr[i] = None
            continue
        if isinstance(lab, int):
            t = torch.full((B,), lab, dtype=torch.int64, device=device)
        else:
            t = lab.detach().to(device, torch.int64).expand(B).contiguous()
            uniform = uniform and bool((t == t[0]).all())
        keep.append(t)
        arr[i] = t.data_ptr()
    return C.cast(arr, C.POINTER(C.c_void_p)), keep, 1 if uniform else 0


def host_coef(rows):
    """[n, m] fp32 HOST table of per-step scalars -> (tensor kept alive, float pointer)."""
    t = torch.as_tensor(rows, dtype=torch.float32).contiguous().cpu()
    return t, C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))


def workspace(device, nbytes):
    return _native.workspace(device, nbytes)


def ptr_array_or_none(tensors):
    if tensors is None:
        return None, []
    arr = (C.c_void_p * len(tensors))()
    keep = []
    for i, t in enumerate(tensors):
        if t is None:
            arr[i] = None
        else:
            t = t.float().contiguous()
            keep.append(t)
            arr[i] = t.data_ptr()
    return C.cast(arr, C.POINTER(C.c_void_p)), keep
