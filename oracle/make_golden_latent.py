"""Golden vectors for the three 2-D latent composition scripts, which are TOP-LEVEL code and cannot be imported:

    shapes/visualize_composition_latent_ito.py    (Ito kappa ODE, "stable" schedule defined in the script)
    shapes/visualize_composition_latent_ito_2.py  (Ito kappa ODE, jax-faithful schedule, clipped kappa)
    mnist/visualize_composition_latent.py         (weighted-sum reverse SDE)

Method: the script's SOURCE is read from /root/reference at generation time and parsed with `ast`; its own function
definitions (schedules, vector_field, get_kappa) and its own sampling statements -- `x = torch.randn(...)`, `dt = ...` and the
`for i in trange(N_STEPS)` loop -- are compiled and executed UNMODIFIED in a namespace where only the surroundings are
stubbed: the reference's MLP class with seeded synthetic weights for the checkpoints, `trange` = a plain range, small N_SAMPLES /
N_STEPS, and a recording noise source for torch.randn / randn_like.  Nothing of the reference is copied into the repo;
only the loop's inputs (initial latents, Hutchinson probes / injected noise) and its output are committed.

Run in the build container:  python -m oracle.make_golden_latent"""
import ast
import sys
import types

import numpy as np
import torch

from . import experts as E
from .make_golden import REF, NoiseTap, _load, _save


def _script_parts(path, fn_names):
    """(function defs named in fn_names, the statements from `x = torch.randn(...)` through the sampling loop)."""
    src = open(path).read()
    tree = ast.parse(src)
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in fn_names]
    assigns = [n for n in tree.body if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], (ast.Name, ast.Tuple))]
    consts = [n for n in assigns if any(isinstance(t, ast.Name) and t.id in ("beta_0", "beta_1", "W1", "W2")
                                        for t in (n.targets[0].elts if isinstance(n.targets[0], ast.Tuple) else [n.targets[0]]))]

    def find_loop(body):
        for i, n in enumerate(body):
            if isinstance(n, ast.For) and isinstance(n.iter, ast.Call) and getattr(n.iter.func, "id", "") == "trange":
                return body, i
            if isinstance(n, ast.With):
                r = find_loop(n.body)
                if r:
                    return r
        return None

    body, li = find_loop(tree.body)
    pre = [n for n in tree.body if body is not tree.body and isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name)
           and n.targets[0].id == "x_gen_history"]      # defined above the `with torch.no_grad():` that holds the loop
    for n in body[:li]:       # the `x = torch.randn(...)` and `dt = ...` statements right before the loop
        if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name) and n.targets[0].id in ("x", "dt", "x_gen_history"):
            pre.append(n)
    mod = ast.Module(body=consts + funcs + pre + [body[li]], type_ignores=[])
    ast.fix_missing_locations(mod)
    return compile(mod, path, "exec")


def _trange(n, **kw):      # tqdm.trange without the progress bar
    return range(n)


def _mlp(ref_mlp, seed):
    m = ref_mlp.MLP(num_out=2).eval()
    m.load_state_dict(E.synth_state_dict(E.mlp_2d_spec(), seed), strict=True)
    return m


def main():
    torch.set_num_threads(4)
    for m in ("matplotlib", "matplotlib.pyplot", "joblib", "tqdm"):
        sys.modules.setdefault(m, types.ModuleType(m))
    N, STEPS = 16, 24
    # ---- shapes: the two Ito scripts -------------------------------------------------------------------------------------
    ref_mlp = _load(f"{REF}/shapes/models/mlp_2d.py", "ref_shapes_mlp")
    jax = _load(f"{REF}/shapes/schedule_jax_faithful.py", "ref_schedule_jax_faithful")
    for name, fn_names, extra in (
        ("visualize_composition_latent_ito", ("stable_log_alpha", "stable_alpha", "stable_sigma", "stable_dlog_alphadt", "stable_beta",
                                              "vector_field", "get_kappa"), {}),
        ("visualize_composition_latent_ito_2", ("vector_field", "get_kappa"),
         dict(alpha=jax.alpha, sigma=jax.sigma, dlog_alphadt=jax.dlog_alphadt, beta=jax.beta)),
    ):
        code = _script_parts(f"{REF}/shapes/{name}.py", fn_names)
        ns = dict(torch=torch, np=np, trange=_trange, DEVICE="cpu", N_SAMPLES=N, N_STEPS=STEPS,
                  shape_model=_mlp(ref_mlp, 401), color_model=_mlp(ref_mlp, 402), **extra)
        with NoiseTap(41) as tap:
            exec(code, ns)
        # draws: x_init, then per step the two Hutchinson probes (shape, colour)
        probes = torch.stack(tap.draws[1:]).view(STEPS, 2, N, 2)
        _save("latent_" + name.replace("visualize_composition_latent_", ""), seed1=401, seed2=402, n_steps=STEPS,
              x_init=tap.draws[0], probes=probes, out=ns["x"].detach())
    # ---- mnist: the weighted-sum SDE script ----------------------------------------------------------------------------------
    ref_mlp_m = _load(f"{REF}/mnist/models/mlp_2d.py", "ref_mnist_mlp")
    sch = _load(f"{REF}/mnist/schedule.py", "ref_mnist_schedule")
    code = _script_parts(f"{REF}/mnist/visualize_composition_latent.py", ())
    ns = dict(torch=torch, np=np, trange=_trange, DEVICE="cpu", N_SAMPLES=N, N_STEPS=STEPS, model1=_mlp(ref_mlp_m, 403),
              model2=_mlp(ref_mlp_m, 404), dlog_alphadt=sch.dlog_alphadt, beta=sch.beta, sigma=sch.sigma, alpha=sch.alpha)
    with NoiseTap(42) as tap, torch.no_grad():       # the script runs this loop under `with torch.no_grad():`
        exec(code, ns)
    _save("latent_sde", seed1=403, seed2=404, n_steps=STEPS, w1=ns["W1"], w2=ns["W2"], x_init=tap.draws[0],
          noise=torch.stack(tap.draws[1:]), out=ns["x"].detach())


if __name__ == "__main__":
    main()
