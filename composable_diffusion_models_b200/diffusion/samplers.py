"""SUPERDIFF sampler (discrete DDPM ancestral step + per-sample Ito log-density estimators).

Drop-in for ``src/diffusion/samplers.py``: ``SuperDiffSampler(sde).sample(model1, model2, batch_size, shape,
device, operation='OR', temp=1.0, bias=0.0)`` and ``.sample_single_model(model, batch_size, shape, device)``.
Per step the experts run, then ONE fused kernel computes the per-sample kappa softmax from the running
log-densities, the combined score, the ancestral update and the K log-density increments (the 2K per-sample
inner products are reduced in-kernel).  Extra keyword ``models=[...]`` runs K > 2 experts.
"""
import ctypes as C

import torch

from .. import _chain, _lib, steps
from ..models import ColoredMNISTScoreModel


class SuperDiffSampler:
    def __init__(self, sde):
        self.sde = sde

    def _scalars(self):
        if hasattr(self.sde, "host_tables"):
            tb = self.sde.host_tables()
        else:
            tb = {k: getattr(self.sde, k).detach().float().cpu() for k in
                  ("betas", "alphas", "sqrt_one_minus_alphas_cumprod", "posterior_variance")}
        return (tb["sqrt_one_minus_alphas_cumprod"].tolist(), tb["betas"].tolist(),
                torch.sqrt(tb["alphas"]).tolist(), torch.sqrt(tb["posterior_variance"]).tolist())

    @torch.no_grad()
    def sample(self, model1, model2, batch_size, shape, device, operation="OR", temp=1.0, bias=0.0, models=None,
               x_init=None, noise=None, seed=None, return_log_q=False, use_chain=None):
        experts = list(models) if models is not None else [model1, model2]
        for m in experts:
            if hasattr(m, "eval"):
                m.eval()
        T = self.sde.num_timesteps
        x = torch.randn((batch_size, *shape), device=device) if x_init is None else x_init.to(device).float().clone()
        log_q = torch.zeros(x.shape[0], len(experts), device=x.device)
        som, betas, sqrt_alpha, sqrt_pv = self._scalars()
        if use_chain is None:
            use_chain = (_chain.native_all(experts, ColoredMNISTScoreModel, x) and x.dim() == 4 and x.shape[2] == x.shape[3]
                         and str(operation).upper() in ("OR", "AND", "AVG"))
        if use_chain:
            x = self._chain(experts, x, log_q, operation, temp, bias, noise, seed, T, som, betas, sqrt_alpha, sqrt_pv)
            out = x.clamp(-1, 1)
            return (out, log_q) if return_log_q else out
        for i in range(T):
            t_idx = T - 1 - i
            t = torch.full((x.shape[0],), t_idx, device=x.device, dtype=torch.long)
            preds = [m(x, t.float()) for m in experts]
            z, rng = None, None
            if i < T - 1:
                if isinstance(noise, str) and noise == "kernel":
                    rng = (seed or 0, i)
                else:
                    z = torch.randn_like(x) if noise is None else (noise(i) if callable(noise) else noise[i]).to(x.device)
            x = steps.step_ddpm_logq(x, preds, log_q, operation, temp, bias, som[t_idx], betas[t_idx],
                                     sqrt_alpha[t_idx], sqrt_pv[t_idx], 1.0 / T, z=z, rng=rng, out=x)
        out = x.clamp(-1, 1)
        return (out, log_q) if return_log_q else out

    def _chain(self, experts, x, log_q, operation, temp, bias, noise, seed, T, som, betas, sqrt_alpha, sqrt_pv):
        """The whole loop through cdm_score_sample_superdiff: ONE host call per chunk of steps (injected / torch-drawn noise is
        staged a chunk at a time in the reference's draw order; in-kernel noise runs the whole chain in one call)."""
        lib = _lib.lib()
        x = x.contiguous()
        B, S, K = x.shape[0], x.shape[2], len(experts)
        if B == 0:
            return x
        rows = [[float(T - 1 - i), som[T - 1 - i], betas[T - 1 - i], sqrt_alpha[T - 1 - i], sqrt_pv[T - 1 - i]] for i in range(T)]
        tab = torch.tensor(rows, dtype=torch.float32)
        harr, hp = _chain.handle_array(experts, x.device)
        op = steps._OPS.get(str(operation).upper(), 2)
        kernel_rng = isinstance(noise, str) and noise == "kernel"
        chunk = T if kernel_rng else max(1, min(T, (256 << 20) // max(1, x.numel() * 4)))
        with torch.cuda.device(x.device):
            prec = _lib.precision_code(experts[0].precision)
            ws = _chain.workspace(x.device, lib.cdm_score_sample_superdiff_workspace_bytes(hp, K, B, S, prec))
            for i0 in range(0, T, chunk):
                m = min(chunk, T - i0)
                z, rng = None, None
                if kernel_rng:
                    rng = C.byref(_lib.Rng(int(seed or 0), i0))
                else:
                    zs = []
                    for i in range(i0, i0 + m):
                        if i < T - 1:       # the last step draws no noise
                            zs.append(torch.randn_like(x) if noise is None else (noise(i) if callable(noise) else noise[i]).to(x.device))
                    if zs:
                        z = torch.stack(zs).float().contiguous()
                ctab, cptr = _chain.host_coef(tab[i0:i0 + m])
                _lib.check(lib.cdm_score_sample_superdiff(hp, K, _lib.ptr(x), _lib.ptr(log_q), op, temp, bias, _lib.ptr(z), rng, cptr, m,
                                                          1 if i0 + m == T else 0, 1.0 / T, B, S, prec, _lib.ptr(ws), ws.numel(),
                                                          _lib.stream_of(x)))
                del ctab
        del harr
        return x

    @torch.no_grad()
    def sample_single_model(self, model, batch_size, shape, device, x_init=None, noise=None, seed=None):
        if hasattr(model, "eval"):
            model.eval()
        T = self.sde.num_timesteps
        x = torch.randn((batch_size, *shape), device=device) if x_init is None else x_init.to(device).float().clone()
        som, betas, sqrt_alpha, sqrt_pv = self._scalars()
        tb_alphas = [a * a for a in sqrt_alpha]
        for i in range(T):
            t_idx = T - 1 - i
            t = torch.full((x.shape[0],), t_idx, device=x.device, dtype=torch.long)
            pred = model(x, t.float())
            z, rng = None, None
            if i < T - 1:
                if isinstance(noise, str) and noise == "kernel":
                    rng = (seed or 0, i)
                else:
                    z = torch.randn_like(x) if noise is None else (noise(i) if callable(noise) else noise[i]).to(x.device)
            # mean = (1/sqrt_alpha)*(x + beta*(-pred/som)) == ancestral form c0*(x - c1*e/c2)
            x = steps.step_cfg(x, [pred], [1.0], 1.0, 1, 1, 1.0 / sqrt_alpha[t_idx], betas[t_idx], som[t_idx],
                               sqrt_pv[t_idx], z=z, rng=rng, out=x)
        del tb_alphas
        return x.clamp(-1, 1)
