"""Reverse-SDE composition of K experts by a weighted SUM of their noise predictions.

Drop-in for ``mnist/compose_scores.py`` (``main(args)`` with ``.model1_path .model2_path .output_file .w1
.w2 .bs .n_steps .xi``) and for the latent loop of ``mnist/visualize_composition_latent.py:63-87``.
Per step the K experts run on libcdm_b200 and ONE fused kernel does combine + Euler-Maruyama update.
"""
import functools
import os

import torch

from . import schedule, steps
from .models import UNet, MLP
from .utils import load_checkpoint


@functools.lru_cache(maxsize=8)
def _sde_coefficients_cached(n_steps, xi):
    return sde_coefficients(n_steps, xi).float().contiguous()


def sde_coefficients(n_steps, xi):
    """Per-step scalars of mnist/compose_scores.py:30-43 for every i at once, fp32, reference op order:
    t, a = dlog_alphadt(t), c = beta(t)/sigma(t), g = sqrt(2*xi*beta(t)) * sqrt(dt).  Returns [n_steps, 4]."""
    dt = 1.0 / n_steps
    t = torch.tensor([1.0 - i * dt for i in range(n_steps)], dtype=torch.float32)
    a = schedule.dlog_alphadt(t)
    c = schedule.beta(t) / schedule.sigma(t)
    g = torch.sqrt(2 * xi * schedule.beta(t)) * torch.sqrt(torch.tensor(dt))
    return torch.stack([t, a, c, g], dim=1).contiguous()


@torch.no_grad()
def sample_composed_sde(experts, weights, bs, shape, n_steps, xi=1.0, device="cuda", x_init=None, noise=None,
                        seed=None, call=None):
    """The hot loop of mnist/compose_scores.py:26-46 for K experts.

    experts[k](x, t) -> eps (``call`` overrides the calling convention, e.g. the MLP's (t, x) order).
    noise: None -> ``torch.randn_like`` per step (the reference's RNG order); a [n_steps, B, ...] tensor or a
    callable i -> tensor -> injected; "kernel" -> drawn inside the fused kernel from (seed, i), zero HBM bytes.
    """
    x = torch.randn(bs, *shape, device=device) if x_init is None else x_init.to(device).float().clone()
    if call is None and _chain_ok(experts, x):
        return _sample_sde_chain(experts, weights, x, n_steps, xi, noise, seed)
    coef = sde_coefficients(n_steps, xi).tolist()
    dt = 1.0 / n_steps
    call = call or (lambda m, xx, tt: m(xx, tt))
    for i in range(n_steps):
        tv, a, c, g = coef[i]
        t = torch.full((x.shape[0],), tv, device=x.device)
        eps = [call(m, x, t) for m in experts]
        if isinstance(noise, str) and noise == "kernel":
            x = steps.step_sde(x, eps, weights, a, c, dt, g, rng=(seed or 0, i), out=x)
        else:
            z = torch.randn_like(x) if noise is None else (noise(i) if callable(noise) else noise[i])
            x = steps.step_sde(x, eps, weights, a, c, dt, g, z=z.to(x.device), out=x)
    return x


def _chain_ok(experts, x):
    """Native unconditional UNet experts of one precision and one channel count: the whole loop runs inside libcdm_b200."""
    if not experts or x.dim() != 4 or x.shape[2] != x.shape[3] or not x.is_cuda:
        return False
    first = experts[0]
    return all(type(m) is UNet and m.num_classes is None and not m.training and m.in_channels == x.shape[1]
               and m.precision == first.precision for m in experts)


def _sample_sde_chain(experts, weights, x, n_steps, xi, noise, seed, steps_per_call=None, step_range=None):
    """sample_composed_sde through cdm_unet_sample_sde: K forwards + the fused step per timestep are enqueued by ONE host
    call per chunk of steps (no Python between kernels), and the time embedding is one row per step and expert.
    Bit-identical to the per-step loop.  Injected / torch-drawn noise is staged ``steps_per_call`` steps at a time (drawn
    in the reference's order); in-kernel noise runs the whole chain in one call.  ``step_range=(i0, i1)`` runs only steps
    i0 .. i1-1 of the n_steps-step chain (noise indices stay absolute), updating ``x`` in place."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    x = x.contiguous()
    B, S = x.shape[0], x.shape[2]
    if B == 0:
        return x
    prec = _lib.precision_code(experts[0].precision)
    coef = _sde_coefficients_cached(n_steps, float(xi))                 # HOST [n_steps, 4]
    handles = (C.c_void_p * len(experts))(*[m._native_handle(x.device).value for m in experts])
    hp = C.cast(handles, C.POINTER(C.c_void_p))
    kernel_rng = isinstance(noise, str) and noise == "kernel"
    if steps_per_call is None:
        steps_per_call = n_steps if kernel_rng else max(1, min(n_steps, (256 << 20) // max(1, x.numel() * 4)))
    from .models import _native
    with torch.cuda.device(x.device):
        ws = _native.workspace(x.device, lib.cdm_unet_sample_workspace_bytes(hp, len(experts), B, S, prec))
        first, last = step_range if step_range is not None else (0, n_steps)
        for i0 in range(first, last, steps_per_call):
            n = min(steps_per_call, last - i0)
            if kernel_rng:
                z, rng = None, C.byref(_lib.Rng(int(seed or 0), i0))
            else:
                rng = None
                if noise is None:
                    z = torch.stack([torch.randn_like(x) for _ in range(n)])
                elif callable(noise):
                    z = torch.stack([noise(i).to(x.device) for i in range(i0, i0 + n)])
                else:
                    z = noise[i0:i0 + n].to(x.device)
                z = z.float().contiguous()
            cf = coef[i0:i0 + n].contiguous()
            _lib.check(lib.cdm_unet_sample_sde(hp, _lib.farray(weights), len(experts), _lib.ptr(x), None, 0, _lib.ptr(z), rng,
                                               C.cast(C.c_void_p(cf.data_ptr()), C.POINTER(C.c_float)), n, 1.0 / n_steps, B, S,
                                               prec, _lib.ptr(ws), ws.numel(), _lib.stream_of(x)))
    return x


_SIDE_STREAMS = {}


@torch.no_grad()
def sample_sde_host_stream(experts, weights, x, n_steps, z_host, x_host=None, xi=1.0, step_range=None):
    """Reverse-SDE steps (mnist/compose_scores.py:30-46) whose injected noise lives in pinned HOST memory and whose
    per-step state is read back to the host: ``z_host[j]`` is the noise of the j-th step of ``step_range`` and
    ``x_host[j]`` (optional) receives x after it.  The copies are pipelined around the compute: the host-to-device copy
    of step j+1's noise and the device-to-host copy of step j's result run on a side stream while step j's kernels run on
    the current stream (two device buffers each way).  ``x`` ([B, C, S, S] on the GPU) is updated in place and returned;
    results are bit-identical to ``sample_composed_sde`` with the same noise.  Native UNet experts only."""
    if not _chain_ok(experts, x):
        raise ValueError("sample_sde_host_stream: needs native unconditional UNet experts in eval mode and a CUDA state")
    first, last = step_range if step_range is not None else (0, n_steps)
    n = last - first
    dev = x.device
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = torch.cuda.Stream(dev)
        zb = [torch.empty_like(x), torch.empty_like(x)]
        xs = [torch.empty_like(x), torch.empty_like(x)] if x_host is not None else None
        h2d = [None, None]      # event: noise of the step using buffer b has landed
        used = [None, None]     # event: the step that read zb[b] has finished
        d2h = [None, None]      # event: the read-back out of xs[b] has finished
        side.wait_stream(main)

        def upload(j):
            b = j & 1
            with torch.cuda.stream(side):
                if used[b] is not None:
                    side.wait_event(used[b])
                zb[b].copy_(z_host[j], non_blocking=True)
                h2d[b] = torch.cuda.Event()
                h2d[b].record(side)

        upload(0)
        for j in range(n):
            b = j & 1
            if j + 1 < n:
                upload(j + 1)
            main.wait_event(h2d[b])
            x = _sample_sde_chain(experts, weights, x, n_steps, xi, (lambda i, t=zb[b]: t), None, step_range=(first + j, first + j + 1))
            used[b] = torch.cuda.Event()
            used[b].record(main)
            if x_host is not None:
                if d2h[b] is not None:
                    main.wait_event(d2h[b])
                xs[b].copy_(x)
                staged = torch.cuda.Event()
                staged.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(staged)
                    x_host[j].copy_(xs[b], non_blocking=True)
                    d2h[b] = torch.cuda.Event()
                    d2h[b].record(side)
        main.wait_stream(side)
    return x


@torch.no_grad()
def sample_composed_latent_sde(experts, weights, n_samples, n_steps, xi=1.0, device="cuda", x_init=None, noise=None,
                               seed=None, precision="fp32"):
    """mnist/visualize_composition_latent.py:63-87 in ONE persistent launch (cdm_mlp_sample_sde): the whole
    n_steps chain of every sample stays on-chip.  experts: native ``MLP`` modules.  ``precision="fp16"`` runs the two
    256x256 hidden layers on tcgen05 (cdm_mlp_sample_sde_tc; K <= 2 experts of the reference's shape)."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    dev = torch.device(device)
    x = torch.randn(n_samples, experts[0].num_out, device=dev) if x_init is None else x_init.to(dev).float().clone()
    x = x.contiguous()
    coef = sde_coefficients(n_steps, xi).to(dev)
    handles = (C.c_void_p * len(experts))(*[m._native_handle(x.device).value for m in experts])
    z = None
    rng = None
    if isinstance(noise, str) and noise == "kernel":
        rng = C.byref(_lib.Rng(int(seed or 0), 0))
    else:
        z = (torch.stack([torch.randn_like(x) for _ in range(n_steps)]) if noise is None else noise).to(dev).float().contiguous()
    fn = lib.cdm_mlp_sample_sde_tc if _lib.precision_code(precision) == _lib.PREC_F16 else lib.cdm_mlp_sample_sde
    with torch.cuda.device(dev):
        _lib.check(fn(C.cast(handles, C.POINTER(C.c_void_p)), _lib.farray(weights), len(experts),
                                          _lib.ptr(x), _lib.ptr(z), rng, _lib.ptr(coef), n_steps, 1.0 / n_steps,
                                          x.shape[0], _lib.stream_of(x)))
    return x


def main(args, x_init=None, noise=None):
    device = "cuda"
    model1 = UNet().to(device).eval()
    load_checkpoint(model1, None, args.model1_path, device)
    model2 = UNet().to(device).eval()
    load_checkpoint(model2, None, args.model2_path, device)
    x = sample_composed_sde([model1, model2], [args.w1, args.w2], args.bs, (1, 28, 28), args.n_steps, args.xi,
                            device=device, x_init=x_init, noise=noise)
    out = getattr(args, "output_file", None)
    if out:
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        try:
            from torchvision.utils import save_image
            save_image(x.clamp(-1, 1), out, nrow=8, normalize=True, value_range=(-1, 1))
        except Exception:   # image writers are not part of the sampler path
            torch.save(x.cpu(), out + ".pt")
    return x
