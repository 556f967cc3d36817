"""GPU: fused combine+update step kernels vs the oracle (teacher-forced, one step at a time).

Tolerances: the elementwise chain is evaluated with explicit round-to-nearest ops in the reference's
operation order, so with injected noise x' must match the fp32 oracle to <= 2 ulp-ish (1e-6 rel-L2);
per-sample reductions (log-q, kappa) differ only in summation order (1e-5)."""
import pytest
import torch

from conftest import rel_l2
from oracle import samplers as OS
from oracle import schedule as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _steps():
    from composable_diffusion_models_b200 import steps
    return steps


@pytest.mark.parametrize("shape,K", [((1, 28, 28), 2), ((1, 28, 28), 1), ((3, 32, 32), 3), ((2,), 2), ((1, 7, 9), 2)])
def test_step_sde(shape, K):
    from composable_diffusion_models_b200.compose_scores import sde_coefficients
    g = torch.Generator().manual_seed(1)
    B, n_steps, xi = 5, 50, 0.8
    x = torch.randn(B, *shape, generator=g)
    eps = [torch.randn(B, *shape, generator=g) for _ in range(K)]
    z = torch.randn(B, *shape, generator=g)
    w = [1.0, 0.7, -0.3][:K]
    coef = sde_coefficients(n_steps, xi).tolist()
    for i in (0, 17, 49):
        tv, a, c, gg = coef[i]
        want = OS.sde_step(x, eps, w, 1.0 - i / n_steps, 1.0 / n_steps, xi, z)
        got = _steps().step_sde(x.to(DEV), [e.to(DEV) for e in eps], w, a, c, 1.0 / n_steps, gg, z=z.to(DEV)).cpu()
        assert rel_l2(got, want) < 1e-6, (shape, K, i)


def test_step_sde_inplace_and_kernel_rng():
    st = _steps()
    x = torch.randn(4, 1, 28, 28, device=DEV)
    e = torch.randn_like(x)
    z = st.fill_normal(x.shape, DEV, (123, 7))
    assert abs(float(z.mean())) < 0.1 and abs(float(z.std()) - 1.0) < 0.1
    a = st.step_sde(x, [e], [1.0], -3.0, 2.0, 1e-3, 0.2, z=z)
    b = st.step_sde(x.clone(), [e], [1.0], -3.0, 2.0, 1e-3, 0.2, rng=(123, 7))
    assert torch.equal(a, b)                      # in-kernel Philox == materialised stream
    xc = x.clone()
    st.step_sde(xc, [e], [1.0], -3.0, 2.0, 1e-3, 0.2, z=z, out=xc)
    assert torch.equal(xc, a)                     # x_out may alias x
    big = st.fill_normal((1 << 20,), DEV, (5, 0))
    assert abs(float(big.mean())) < 5e-3 and abs(float(big.std()) - 1.0) < 5e-3
    assert abs(float((big ** 4).mean()) - 3.0) < 0.05


@pytest.mark.parametrize("S", [16, 64])
def test_step_ddim_and_gray(S):
    g = torch.Generator().manual_seed(2)
    B = 3
    x = torch.randn(B, 3, S, S, generator=g)
    es = torch.randn(B, 1, S, S, generator=g)
    ec = torch.randn(B, 3, S, S, generator=g)
    ts = OS.ddim_time_grid(10)
    al, sg = S_alpha(ts), S_sigma(ts)
    for i in (0, 9):
        want = OS.ddim_step(x, es, ec, 1.0, 0.6, ts[i], ts[i + 1])
        gray = torch.empty(B, 1, S, S, device=DEV)
        got = _steps().step_ddim(x.to(DEV), [es.to(DEV), ec.to(DEV)], [1.0, 0.6], 1.0 + 0.6, float(al[i]), float(sg[i]),
                                 float(al[i + 1]), float(sg[i + 1]), gray_out=gray)
        assert rel_l2(got.cpu(), want) < 1e-6
        assert rel_l2(gray.cpu(), OS.grayscale(want)) < 1e-6
    assert rel_l2(_steps().grayscale(x.to(DEV)).cpu(), OS.grayscale(x)) < 1e-7


def S_alpha(t):
    return S.alpha(t)


def S_sigma(t):
    return S.sigma(t)


@pytest.mark.parametrize("op", ["OR", "AND", "AVG"])
@pytest.mark.parametrize("K", [2, 4])
@pytest.mark.parametrize("size", [32, 64])      # 64x64 with a small batch: a cluster of CTAs shares each sample (DSMEM reduction)
def test_step_ddpm_logq(op, K, size):
    g = torch.Generator().manual_seed(3)
    B, shape = 6, (3, size, size)
    sde = S.VPSDETables(num_timesteps=100)
    x = torch.randn(B, *shape, generator=g)
    ns = [torch.randn(B, *shape, generator=g) for _ in range(K)]
    z = torch.randn(B, *shape, generator=g)
    logq = torch.randn(B, K, generator=g) * 3
    for t_idx, last in ((99, False), (40, False), (0, True)):
        if K == 2:
            want_x, want_q = OS.superdiff_step(sde, x, ns, logq, t_idx, z, op, 1.5, 0.2, last=last)
        else:
            want_x, want_q = OS.superdiff_step(sde, x, ns, logq, t_idx, z, op if op != "AVG" else "OR", 1.5, 0.2, last=last)
        q = logq.clone().to(DEV)
        kap = torch.empty(B, K, device=DEV)
        opk = op if (K == 2 or op != "AVG") else "OR"
        got = _steps().step_ddpm_logq(x.to(DEV), [n.to(DEV) for n in ns], q, opk, 1.5, 0.2,
                                      float(sde.sqrt_one_minus_alphas_cumprod[t_idx]), float(sde.betas[t_idx]),
                                      float(torch.sqrt(sde.alphas[t_idx])), float(torch.sqrt(sde.posterior_variance[t_idx])),
                                      1.0 / 100, z=None if last else z.to(DEV), kappa_out=kap)
        assert rel_l2(got.cpu(), want_x) < 1e-6
        assert rel_l2(q.cpu(), want_q) < 1e-5
        assert rel_l2(kap.cpu(), OS.superdiff_kappas(logq, opk, 1.5, 0.2)) < 1e-6


@pytest.mark.parametrize("variant", ["beta", "g2"])
def test_step_ode_kappa_image(variant):
    g = torch.Generator().manual_seed(4)
    B, S_ = 4, 16
    x = torch.randn(B, 3, S_, S_, generator=g)
    es = torch.randn(B, 1, S_, S_, generator=g)
    ec = torch.randn(B, 3, S_, S_, generator=g)
    d1, d2 = torch.randn(B, generator=g) * 50, torch.randn(B, generator=g) * 50
    for t_val in (1.0, 0.37, 0.004):
        dt = 1e-3
        t = torch.full((B,), t_val)
        scale = 3.0 if variant == "beta" else 1.0
        want_x, want_k = OS.ito_ode_step(x, es.repeat(1, 3, 1, 1), ec, scale * d1, d2, t_val, dt, variant)
        coef = 0.5 * (S.beta(t) if variant == "beta" else S.g2(t))
        kap = torch.empty(B, device=DEV)
        got = _steps().step_ode_kappa(x.to(DEV), es.to(DEV), ec.to(DEV), d1.to(DEV), d2.to(DEV), float(S.sigma(t)[0]),
                                      float(S.dlog_alphadt(t)[0]), float(coef[0]), dt, mode=0, div1_scale=scale,
                                      kappa_out=kap)
        assert rel_l2(kap.cpu(), want_k) < 2e-5
        assert rel_l2(got.cpu(), want_x) < 1e-5


@pytest.mark.parametrize("variant,mode", [("stable", 2), ("clipped", 1)])
def test_step_ode_kappa_latent(variant, mode):
    g = torch.Generator().manual_seed(5)
    B = 64
    x = torch.randn(B, 2, generator=g)
    e1, e2 = torch.randn(B, 2, generator=g), torch.randn(B, 2, generator=g)
    d1, d2 = torch.randn(B, generator=g), torch.randn(B, generator=g)
    for t_val in (0.9, 0.2):
        t = torch.full((B,), t_val)
        want_x, want_k = OS.latent_ito_step(x, e1, e2, d1, d2, t_val, 1e-3, variant)
        if variant == "stable":
            sig, coef, den = S.stable_sigma(t)[0], S.stable_beta(t)[0], 1e-9
        else:
            sig, coef, den = S.jax_sigma(t)[0], S.jax_beta(t)[0], 1e-5
        kap = torch.empty(B, device=DEV)
        got = _steps().step_ode_kappa(x.to(DEV), e1.to(DEV), e2.to(DEV), d1.to(DEV), d2.to(DEV), float(sig),
                                      float(S.dlog_alphadt(t)[0]), float(coef), 1e-3, mode=mode, den_eps=den, kappa_out=kap)
        assert rel_l2(kap.cpu(), want_k) < 1e-5
        assert rel_l2(got.cpu(), want_x) < 1e-6


def test_step_cfg_both_forms():
    g = torch.Generator().manual_seed(6)
    B, shape = 3, (3, 32, 32)
    x = torch.randn(B, *shape, generator=g)
    pu, ps, pc = (torch.randn(B, *shape, generator=g) for _ in range(3))
    z = torch.randn(B, *shape, generator=g)
    acp = S.ddpm_alphas_cumprod(500)
    ab = acp[123]
    want = OS.cfg_x0_step(ps, pc, pu, 7.5, 7.5, ab)
    got = _steps().step_cfg(x.to(DEV), [pu.to(DEV), ps.to(DEV), pc.to(DEV)], [1.0, 7.5, 7.5], 1.0, 0, 0,
                            float(torch.sqrt(ab)), float(torch.sqrt(1.0 - ab)))
    assert rel_l2(got.cpu(), want) < 1e-6
    sde = S.VPSDETables(num_timesteps=200)
    i = 77
    sra = torch.sqrt(1.0 / sde.alphas)[i]
    want = OS.weighted_ddpm_step(x, [ps, pc, pu], [1.0, 2.0, 0.5], sde.betas[i], sde.sqrt_one_minus_alphas_cumprod[i],
                                 sra, sde.posterior_variance[i], z)
    got = _steps().step_cfg(x.to(DEV), [ps.to(DEV), pc.to(DEV), pu.to(DEV)], [1.0, 2.0, 0.5], 3.5, 1, 1, float(sra),
                            float(sde.betas[i]), float(sde.sqrt_one_minus_alphas_cumprod[i]),
                            float(torch.sqrt(sde.posterior_variance[i])), z=z.to(DEV))
    assert rel_l2(got.cpu(), want) < 1e-6


def test_step_argument_errors():
    st = _steps()
    x = torch.randn(2, 3, 8, 8, device=DEV)
    with pytest.raises(ValueError):
        st.step_sde(x, [torch.randn(2, 2, 8, 8, device=DEV)], [1.0], 0.0, 0.0, 0.1, 0.1, z=x)   # channels not 1 or C
    with pytest.raises(ValueError):
        st.step_sde(x, [x], [1.0], 0.0, 0.0, 0.1, 0.1)                                            # neither z nor rng
    from composable_diffusion_models_b200 import _lib
    with pytest.raises(_lib.CdmError):
        st.step_sde(x.cpu(), [x.cpu()], [1.0], 0.0, 0.0, 0.1, 0.1, z=x.cpu())                     # no CPU path
    empty = torch.empty(0, 3, 8, 8, device=DEV)
    assert st.step_sde(empty, [empty], [1.0], 0.0, 0.0, 0.1, 0.1, z=empty).shape[0] == 0          # empty batch is a no-op


@pytest.mark.parametrize("B,L,D", [(512, 2, 784), (33, 2, 12288), (7, 5, 64), (1, 8, 4)])
def test_latent_decode_vs_numpy(B, L, D):
    """PCA inverse transform (section 8(f) row 3) vs the reference's numpy expression, and vs sklearn when L fits."""
    from composable_diffusion_models_b200 import steps
    from oracle import samplers as OS
    g = torch.Generator().manual_seed(B + D)
    z = torch.randn(B, L, generator=g)
    comp = torch.randn(L, D, generator=g) / D ** 0.5
    mean = torch.randn(D, generator=g)
    want = OS.pca_decode(z, comp.numpy(), mean.numpy())
    got = steps.decode_latents(z.to(DEV), comp, mean).cpu()
    assert rel_l2(got, want) < 1e-6
    if B >= 16:
        from sklearn.decomposition import PCA
        pca = PCA(n_components=L)
        data = torch.randn(64, D, generator=g).numpy()
        pca.fit(data)
        got = steps.decode_latents(z.to(DEV), pca.components_, pca.mean_).cpu()
        assert rel_l2(got, torch.from_numpy(pca.inverse_transform(z.numpy())).float()) < 1e-5


def test_cluster_split_is_invisible(monkeypatch):
    """The step kernels give a small batch of large samples to clusters of CTAs (one sample per cluster, reductions through
    distributed shared memory).  Elementwise outputs must not depend on the split at all; reductions only in summation order."""
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, ".")
from composable_diffusion_models_b200 import steps
g = torch.Generator().manual_seed(1)
B = 5
x = torch.randn(B, 3, 64, 64, generator=g).cuda()
ns = [torch.randn(B, 3, 64, 64, generator=g).cuda() for _ in range(3)]
z = torch.randn(B, 3, 64, 64, generator=g).cuda()
lq = (torch.randn(B, 3, generator=g) * 2).cuda()
out = steps.step_ddpm_logq(x, ns, lq, "OR", 1.2, 0.1, 0.9, 0.01, 0.99, 0.05, 1e-3, z=z)
e1 = torch.randn(B, 1, 64, 64, generator=g).cuda()
d1, d2 = torch.randn(B, generator=g).cuda() * 30, torch.randn(B, generator=g).cuda() * 30
ok = steps.step_ode_kappa(x, e1, ns[0], d1, d2, 0.9, -5.0, 2.0, 1e-3, div1_scale=3.0)
gray = torch.empty(B, 1, 64, 64, device="cuda")
dd = steps.step_ddim(x, [e1, ns[1]], [1.0, 1.0], 2.0, 0.5, 0.8, 0.6, 0.7, gray_out=gray)
torch.save(dict(out=out.cpu(), lq=lq.cpu(), ok=ok.cpu(), dd=dd.cpu(), gray=gray.cpu()), sys.argv[1])
'''
    import os
    import tempfile
    res = {}
    for split in ("1", "4", "8"):
        f = os.path.join(tempfile.mkdtemp(), "o.pt")
        env = dict(os.environ, CDM_STEP_SPLIT=split)
        subprocess.run([sys.executable, "-c", code, f], check=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        res[split] = torch.load(f)
    for split in ("4", "8"):
        assert torch.equal(res[split]["dd"], res["1"]["dd"]) and torch.equal(res[split]["gray"], res["1"]["gray"])
        assert rel_l2(res[split]["out"], res["1"]["out"]) < 1e-30            # the update is elementwise given kappa
        assert rel_l2(res[split]["lq"], res["1"]["lq"]) < 1e-5
        assert rel_l2(res[split]["ok"], res["1"]["ok"]) < 1e-5
