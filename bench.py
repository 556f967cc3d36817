#!/usr/bin/env python
"""Benchmark: composed samples/sec of the K-expert composed-score sampler on B200.

Workload (BASELINE.json configs[1]): MNIST 1x28x28 UNet experts, K=2 weighted-sum
composition (mnist/compose_scores.py), reverse-SDE update, 1000-step chain, batch 4096 per GPU.
One bench "step" = one sampler timestep over the whole batch = K UNet forwards + one fused
combine/update launch.  A sample needs CHAIN_STEPS=1000 such steps, so

    composed samples/sec = n_gpus * B / (CHAIN_STEPS * seconds_per_step)

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]; under torchrun each rank
drives one GPU with its own independent chains (weak scaling, no collective in the loop); rank 0 prints
ONE JSON line.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks.  Besides the contract keys the line carries (N = 1 only, each leg a few seconds):

  parity              rel-L2 of final samples vs the fp32 CPU oracle on the bench workload at small batch: this
                      library's fp16 / f16x3 / fp32 modes AND the same torch ops on CUDA with cuDNN TF32 (the
                      reference's real GPU arithmetic); plus run-to-run bit-identity of the benchmarked mode
  gpu_eager_baseline  the reference's own modules run eagerly on this B200 (cuDNN TF32): the GPU path to beat
  full_chain_s        one real 1000-step chain, wall time on the device (no extrapolation)
  modes               ms/step of the other precision modes of this library on the same workload
  configs             C1 / C3 / C4 / C5 of BASELINE.json at their named sizes, a few timed steps each
  strong              the strong-scaling reading of the workload: 4096 samples in TOTAL (4096 / N per GPU)
  grouped             K = 2 expert forwards as ONE grouped launch per convolution vs back to back, at small batches

See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHAIN_STEPS = 1000
BATCH = 4096
K_EXPERTS = 2
IMG = (1, 28, 28)
UNET_GFLOP = 0.797447   # per sample per forward, sum over the 13 tensor-core GEMMs + init/out convs
METRIC = "composed samples/sec (K=2 expert MNIST UNet reverse-SDE sampling, 1000 steps)"
WORKLOAD = "mnist_unet_K2_sde_1000steps_B4096"
CPU_BATCH = 256         # samples per step on the CPU arm (64-256; cost is linear in B)


def _traffic():
    """DRAM bytes per conv launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the tensor-core conv
    launches of one expert forward at B=4096) from the committed `ncu --set full` capture, or None."""
    for name in ("r02_conv_dram_traffic.json", "r01_conv_dram_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["avg_bytes_per_launch"]
        except Exception:
            continue
    return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------
# The reference's own implementation of the path (mnist/compose_scores.py:26-46): its UNet module and schedule
# functions staged under oracle/_ref by oracle/build_ref.py (kind "reference"), else the oracle port (kind "port").
# ---------------------------------------------------------------------------------------------------------
class ReferenceSampler:
    """K = 2 weighted-sum reverse-SDE steps exactly as mnist/compose_scores.py:30-46 writes them, on `device`."""

    def __init__(self, device, channels_last=False):
        import torch
        from oracle import build_ref
        self.torch, self.device = torch, torch.device(device)
        ref = build_ref.load_mnist()
        self.kind = "reference" if ref is not None else "port"
        if ref is not None:
            RefUNet, self.sched = ref
            self.models = []
            for k in range(K_EXPERTS):
                torch.manual_seed(1234 + k)                   # the same synthetic experts as the CUDA arm
                m = RefUNet().to(self.device).eval()
                self.models.append(m.to(memory_format=torch.channels_last) if channels_last else m)
            self.forward = lambda k, x, t: self.models[k](x, t)
        else:
            from oracle import experts as E
            from oracle import schedule as S
            self.sched = S
            sds = [{n: v.to(self.device) for n, v in E.synth_state_dict(E.unet_small_spec(1), 1234 + k).items()}
                   for k in range(K_EXPERTS)]
            self.forward = lambda k, x, t: E.unet_small_forward(sds[k], x, t)
        self.channels_last = channels_last

    def step(self, x, i, dt, xi=1.0, w=(1.0, 1.0)):
        torch, s = self.torch, self.sched
        t_val = 1.0 - i * dt
        t = torch.full((x.shape[0],), t_val, device=self.device)
        e = w[0] * self.forward(0, x, t) + w[1] * self.forward(1, x, t)
        drift = s.dlog_alphadt(t).view(-1, 1, 1, 1) * x - s.beta(t).view(-1, 1, 1, 1) / s.sigma(t).view(-1, 1, 1, 1) * e
        diffusion = torch.sqrt(2 * xi * s.beta(t)).view(-1, 1, 1, 1)
        return x + (-drift * dt + diffusion * torch.sqrt(torch.tensor(dt)) * torch.randn_like(x))


def reference_cpu_step_seconds(batch, steps, warmup, threads):
    """Seconds per sampler step of the reference path on the host cores (fp32, torch CPU ops, all threads)."""
    import torch
    torch.set_num_threads(threads)
    rs = ReferenceSampler("cpu")
    x = torch.randn(batch, *IMG, generator=torch.Generator().manual_seed(0))
    dt = 1.0 / CHAIN_STEPS
    with torch.no_grad():
        for i in range(warmup):
            x = rs.step(x, i, dt)
        t0 = time.perf_counter()
        for i in range(steps):
            x = rs.step(x, warmup + i, dt)
        sec = (time.perf_counter() - t0) / steps
    return sec, rs.kind


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    sec, kind = reference_cpu_step_seconds(CPU_BATCH, steps, warmup, threads)
    value = CPU_BATCH / (CHAIN_STEPS * sec)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_timed": CPU_BATCH, "chain_steps": CHAIN_STEPS},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": kind,
                         "sample": f"{steps} sampler steps at batch {CPU_BATCH} (of the 1000-step, batch-4096 workload), fp32, "
                                   + ("the reference's own UNet + schedule modules (oracle/_ref)" if kind == "reference"
                                      else "oracle port (oracle/_ref not staged)") + ", torch CPU ops"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# extra legs of the CUDA arm (N = 1)
# ---------------------------------------------------------------------------------------------------------
def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _timed(fn, steps, warm):
    import torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def parity_block(dev, modes, n_steps=CHAIN_STEPS, B=2):
    """Final-sample rel-L2 vs the fp32 CPU oracle on the bench workload (the bench's own synthetic experts, weights (1, 1),
    injected noise) at small batch: this library's modes, and the same torch ops on CUDA with cuDNN TF32."""
    import torch
    from composable_diffusion_models_b200.compose_scores import sample_composed_sde
    from composable_diffusion_models_b200.models import UNet
    from oracle import experts as E
    from oracle import samplers as OS
    sds = []
    for k in range(K_EXPERTS):
        torch.manual_seed(1234 + k)
        sds.append({n: v.detach().clone() for n, v in UNet().state_dict().items()})
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(B, *IMG, generator=g)
    noise = torch.randn(n_steps, B, *IMG, generator=g)
    w = [1.0] * K_EXPERTS
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        want = OS.sample_sde([lambda x, t, sd=sd: E.unet_small_forward(sd, x, t) for sd in sds], w, x0, noise, n_steps, 1.0)
    out = {"workload": f"{WORKLOAD} at batch {B}: the whole {n_steps}-step chain (dt = 1/{n_steps}), weights (1, 1), injected noise",
           "metric": "relative L2 of the final samples vs the fp32 CPU oracle", "tolerance_bf16_tf32": 1e-3, "tolerance_fp32": 1e-5}

    def ours(prec):
        experts = []
        for sd in sds:
            m = UNet(precision=prec)
            m.load_state_dict(sd, strict=True)
            experts.append(m.to(dev).eval())
        return sample_composed_sde(experts, w, B, IMG, n_steps, 1.0, device=dev, x_init=x0, noise=noise).cpu()

    for prec in modes:
        try:
            a = ours(prec)
            out[prec] = _rel_l2(a, want)
            if prec == modes[0]:
                out["deterministic"] = bool(torch.equal(a, ours(prec)))      # two runs of the benchmarked mode, bit for bit
        except Exception as e:   # a mode this build lacks is reported, not fatal
            out[prec] = f"error: {e}"
    # the reference's real GPU arithmetic: the same functional torch ops on CUDA tensors, conv2d in TF32 (torch default)
    torch.backends.cudnn.allow_tf32 = True
    csds = [{n: v.to(dev) for n, v in sd.items()} for sd in sds]
    with torch.no_grad():
        got = OS.sample_sde([lambda x, t, sd=sd: E.unet_small_forward(sd, x.to(dev), t.to(dev)).cpu() for sd in csds], w, x0, noise,
                            n_steps, 1.0)
    out["torch_cuda_tf32"] = _rel_l2(got, want)
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        got = OS.sample_sde([lambda x, t, sd=sd: E.unet_small_forward(sd, x.to(dev), t.to(dev)).cpu() for sd in csds], w, x0, noise,
                            n_steps, 1.0)
    out["torch_cuda_fp32"] = _rel_l2(got, want)
    torch.backends.cudnn.allow_tf32 = True
    return out


def gpu_eager_baseline(dev, B, steps=5, warm=2):
    """The reference's own modules, eager PyTorch on this GPU (cuDNN TF32 convs as torch ships): SURVEY.md 8(d)."""
    import torch
    torch.backends.cudnn.allow_tf32 = True
    res = {}
    for cl in (False, True):
        rs = ReferenceSampler(dev, channels_last=cl)
        x = torch.randn(B, *IMG, device=dev)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        dt = 1.0 / CHAIN_STEPS
        state = {"x": x}

        def one(i):
            with torch.no_grad():
                state["x"] = rs.step(state["x"], i, dt)
        ms = _timed(one, steps, warm)
        res["channels_last" if cl else "nchw"] = ms
        kind = rs.kind
        del rs, state, x
        torch.cuda.empty_cache()
    best = min(res.values())
    return {"value": B / (CHAIN_STEPS * best * 1e-3), "unit": "samples/s", "ms_per_step": best, "ms_per_step_by_layout": res,
            "kind": kind, "arithmetic": "eager torch on CUDA, cuDNN TF32 convs (torch default), fp32 elsewhere", "batch": B,
            "tflops": K_EXPERTS * B * UNET_GFLOP / best}


def other_configs(dev, pk):
    """C1 / C3 / C4 / C5 of BASELINE.json at their named sizes: a few timed steps each (synthetic experts)."""
    import torch
    from composable_diffusion_models_b200 import steps as S, schedule
    from composable_diffusion_models_b200.compose_images_ddim import ddim_tables
    from composable_diffusion_models_b200.compose_scores import sample_composed_latent_sde
    from composable_diffusion_models_b200.models import UNet, GuidedUNet, MLP
    out = {}

    def row(name, ms, B, chain, gflop_per_sample_step, **kw):
        tf = B * gflop_per_sample_step / ms
        out[name] = dict(batch=B, chain_steps=chain, ms_per_step=ms, samples_per_s=B / (chain * ms * 1e-3), tflops=tf,
                         frac_of_tensor_peak=tf / pk["tf_sust"], **kw)

    def guard(name, fn):
        try:
            fn()
        except Exception as e:
            out[name] = {"error": str(e)[:300]}
        torch.cuda.empty_cache()

    def c1():
        B = 1 << 20
        torch.manual_seed(0)
        experts = [MLP().to(dev).eval() for _ in range(2)]
        sample_composed_latent_sde(experts, [1.0, 1.0], 4096, 50, device=dev, noise="kernel", seed=1, precision="fp16")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sample_composed_latent_sde(experts, [1.0, 1.0], B, 1000, device=dev, noise="kernel", seed=2, precision="fp16")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 1000
        row("C1", ms, B, 1000, 2 * 0.265e-3, workload="2-D latent MLP experts, K=2 SDE, whole 1000-step chain in one persistent launch (fp16 tcgen05)")

    def c3():
        B = 8192
        torch.manual_seed(0)
        ms_ = UNet(in_channels=1, num_classes=3).to(dev).eval()
        mc_ = UNet(in_channels=3, num_classes=3).to(dev).eval()
        st = {"x": torch.randn(B, 3, 64, 64, device=dev)}
        xg = S.grayscale(st["x"])
        sl, cl = torch.full((B,), 2, device=dev), torch.full((B,), 1, device=dev)
        ts, al, sg = [v.tolist() for v in ddim_tables(50)]
        tv = torch.empty(B, device=dev)

        def one(i):
            i %= 50
            tv.fill_(ts[i])
            es, ec = ms_(xg, tv, sl), mc_(st["x"], tv, cl)
            S.step_ddim(st["x"], [es, ec], [1.0, 1.0], 2.0, al[i], sg[i], al[i + 1], sg[i + 1], out=st["x"], gray_out=xg)
        row("C3", _timed(one, 4, 2), B, 50, 4.166 + 4.177, workload="shapes 64x64 two-expert DDIM (compose_images_ddim), fp16 tcgen05")

    def c4():
        from composable_diffusion_models_b200 import compose_images_ito as ITO
        B = 1024
        if hasattr(ITO, "bench_step_k4"):          # K = 4 Ito superposition (2 shape-type + 2 colour-type experts)
            fn, K = ITO.bench_step_k4(dev, B), 4
        else:
            fn, K = None, 2
        if fn is None:
            torch.manual_seed(0)
            ms_ = UNet(in_channels=1, num_classes=3).to(dev).eval()
            mc_ = UNet(in_channels=3, num_classes=3).to(dev).eval()
            st = {"x": torch.randn(B, 3, 64, 64, device=dev)}
            sl, cl = torch.full((B,), 2, device=dev), torch.full((B,), 1, device=dev)
            tv = torch.empty(B, device=dev)

            def fn(i):
                t_val = 1.0 - (i % 1000) * 1e-3
                tv.fill_(t_val)
                xg = S.grayscale(st["x"])
                es, ds = ms_.forward_jvp(xg, tv, sl, torch.randn_like(xg))
                ec, dc = mc_.forward_jvp(st["x"], tv, cl, torch.randn_like(st["x"]))
                tt = torch.tensor(t_val)
                S.step_ode_kappa(st["x"], es, ec, ds, dc, float(schedule.sigma(tt)), float(schedule.dlog_alphadt(tt)),
                                 0.5 * float(schedule.beta(tt)), 1e-3, mode=0, div1_scale=3.0, out=st["x"])
        row("C4", _timed(fn, 3, 2), B, 1000, K * (4.166 + 4.177), experts=K,
            workload=f"shapes 64x64 Ito kappa-ODE superposition, K={K} experts, Hutchinson divergence by forward-mode JVP "
                     "(primal + tangent through every expert each step), fp16 tcgen05")

    def c5():
        B = 2048
        torch.manual_seed(0)
        g = GuidedUNet(precision="fp16").to(dev).eval()
        st = {"x": torch.randn(B, 3, 32, 32, device=dev)}
        tt = torch.empty(B, device=dev)
        d, c = torch.full((B,), 7, device=dev), torch.full((B,), 2, device=dev)
        nd, nc = torch.full((B,), 10, device=dev), torch.full((B,), 3, device=dev)

        def one(i):
            tt.fill_(float(499 - i))
            pu, ps, pc = g(st["x"], tt, nd, nc), g(st["x"], tt, d, nc), g(st["x"], tt, nd, c)
            S.step_cfg(st["x"], [pu, ps, pc], [1.0, 7.5, 7.5], 1.0, 0, 0, 0.9, 0.4, out=st["x"])
        row("C5", _timed(one, 3, 2), B, 500, 3 * 2.482, workload="colored-MNIST GuidedUNet (cross-attention) CFG, 3 forwards/step, 2048 per GPU, fp16 tcgen05")

    def c4s():
        # C4': the SuperDiff (log-density softmax) flavour of config 4 on the reference's BatchNorm score UNets, K = 4
        from composable_diffusion_models_b200.diffusion import SuperDiffSampler
        from composable_diffusion_models_b200.models import ColoredMNISTScoreModel
        from composable_diffusion_models_b200.schedule import VPSDE
        B, T = 1024, 8
        torch.manual_seed(0)
        experts = [ColoredMNISTScoreModel(precision="fp16").to(dev).eval() for _ in range(4)]
        sampler = SuperDiffSampler(VPSDE(num_timesteps=T, device=dev))
        x0 = torch.randn(B, 3, 32, 32, device=dev)

        def chain(i):      # T steps per call through the whole-chain entry (cdm_score_sample_superdiff)
            sampler.sample(experts[0], experts[1], B, (3, 32, 32), dev, operation="OR", models=experts, x_init=x0, noise="kernel", seed=i)
        ms = _timed(chain, 3, 2) / T
        row("C4_superdiff", ms, B, 1000, 4 * 0.663, experts=4,
            workload="SuperDiff OR, K=4 BatchNorm score UNets 3x32x32 (4x4 strided / transposed convs), fp16 tcgen05, whole-chain entry")

    def su():
        from composable_diffusion_models_b200.models import SimpleUnet
        B = 256
        torch.manual_seed(0)
        m = SimpleUnet(3, precision="fp16").to(dev).eval()
        x = torch.randn(B, 3, 64, 64, device=dev)
        tt, yy = torch.full((B,), 250.0, device=dev), torch.full((B,), 1, device=dev)
        ms = _timed(lambda i: m(x, tt, yy), 3, 2)
        row("simple_unet", ms, B, 1, 11.46, workload="62 M-parameter SimpleUnet, ONE forward at 3x64x64, fp16 tcgen05")

    for name, fn in (("C1", c1), ("C3", c3), ("C4", c4), ("C4_superdiff", c4s), ("C5", c5), ("simple_unet", su)):
        guard(name, fn)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="samples per GPU (default: the named workload's 4096)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "f16x3", "fp32"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the batch is 4096 in TOTAL (4096 / N per GPU) instead of 4096 per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / eager / full-chain / other-config legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from composable_diffusion_models_b200 import _lib
    from composable_diffusion_models_b200.models import UNet

    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().cdm_device_check(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # synthetic experts: the reference architecture, default init under fixed seeds
    def make_experts(prec):
        ex = []
        for k in range(K_EXPERTS):
            torch.manual_seed(1234 + k)
            ex.append(UNet(precision=prec).to(dev).eval())
        return ex

    experts = make_experts(args.precision)
    B = args.batch if args.scaling == "weak" else max(1, args.batch // world)
    gen = torch.Generator(device="cpu").manual_seed(rank)
    dt = 1.0 / CHAIN_STEPS
    weights = [1.0] * K_EXPERTS

    from composable_diffusion_models_b200.compose_scores import _sample_sde_chain

    def run_steps(ex, i0, n, x, z=None):
        """steps i0 .. i0+n-1 of the 1000-step chain through the public sampler path (cdm_unet_sample_sde: K forwards +
        the fused step per timestep, one host call); z = that step's injected noise, else in-kernel Philox (rank, i)."""
        i0 %= CHAIN_STEPS
        n = min(n, CHAIN_STEPS - i0)
        noise = "kernel" if z is None else (lambda i: z)
        return _sample_sde_chain(ex, weights, x, CHAIN_STEPS, 1.0, noise, rank, step_range=(i0, i0 + n))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            tm = torch.tensor([ms], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
        return ms

    def device_resident(ex, Bx, n_steps, n_warm):
        """ms per step of n_steps timed steps (one host call, kernels back to back), inputs resident; max over ranks."""
        x = torch.randn(Bx, *IMG, generator=gen).to(dev)
        for i in range(n_warm):
            x = run_steps(ex, i, 1, x)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        x = run_steps(ex, n_warm, n_steps, x)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n_steps, _lib.launch_count() - l0, x

    # ---- device-resident throughput (`value`) -----------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_per_step, launches, x = device_resident(experts, B, args.steps, warmup)
    sampler.stop_flag = True
    value = world * B / (CHAIN_STEPS * ms_per_step * 1e-3)
    state_absmax = float(x.abs().max())       # the chain state must stay finite (random-init experts make it grow)
    if not (state_absmax < float("inf")):
        raise RuntimeError("bench: the sampler state is not finite")

    # ---- end-to-end through the public API with host buffers (`e2e`) ------------------------------
    # every step: H2D of that step's injected noise from pinned memory, D2H of the step's result -- through
    # compose_scores.sample_sde_host_stream, which pipelines the copies of neighbouring steps around the compute
    from composable_diffusion_models_b200.compose_scores import sample_sde_host_stream
    z_host = torch.randn(B, *IMG, generator=gen).pin_memory()
    x_host = torch.empty(B, *IMG).pin_memory()
    x = sample_sde_host_stream(experts, weights, x, CHAIN_STEPS, [z_host] * 2, [x_host] * 2, step_range=(0, 2))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x = sample_sde_host_stream(experts, weights, x, CHAIN_STEPS, [z_host] * args.steps, [x_host] * args.steps, step_range=(0, args.steps))
    e1.record()
    barrier()
    ms2 = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B / (CHAIN_STEPS * (ms2 / args.steps) * 1e-3)
    step_bytes = B * IMG[0] * IMG[1] * IMG[2] * 4

    # ---- strong-scaling reading of the workload: 4096 samples in total -----------------------------
    if args.scaling == "strong" or world == 1:
        strong = {"batch_total": B * world, "batch_per_gpu": B, "ms_per_step": ms_per_step, "value": value, "unit": "samples/s"}
    else:
        Bs = max(1, BATCH // world)
        ms_s, _, _ = device_resident(experts, Bs, args.steps, warmup)
        strong = {"batch_total": Bs * world, "batch_per_gpu": Bs, "ms_per_step": ms_s, "unit": "samples/s",
                  "value": world * Bs / (CHAIN_STEPS * ms_s * 1e-3)}

    # ---- per-kernel-class timing (roofline) on rank 0: a separate short pass, events around every launch
    roof, classes = None, None
    pk = _peaks()
    if rank == 0:
        nprof = min(args.steps, 10)
        for i in range(2):                      # back to steady state after the copy-bound e2e leg
            x = run_steps(experts, i, 1, x)
        torch.cuda.synchronize()
        _lib.prof_enable(True)
        for i in range(nprof):
            x = run_steps(experts, i, 1, x)
        torch.cuda.synchronize()
        classes = _lib.prof_summary()
        _lib.prof_enable(False)
        total_ms = sum(c["ms"] for c in classes.values())
        conv = classes.get("conv_tc") or classes.get("conv_fp32")
        if conv:
            ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
            roof = {"kernel": ("conv_halo_kernel / conv_stack3_kernel / conv_tc_kernel (tcgen05 implicit-GEMM 3x3 convs, all launches)"
                               if "conv_tc" in classes else "conv_fp32_kernel"),
                    "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                    "traffic": _traffic() if "conv_tc" in classes and B == BATCH and args.precision == "fp16" else None,
                    "peak_source": pk["src"] + " (sustained 16-bit cuBLAS; fp16 and bf16 share the tcgen05 kind::f16 rate)",
                    "share_of_step": conv["ms"] / total_ms, "avg_launch_ms": conv["ms"] / conv["launches"]}
        st = classes.get("step")
        if st and roof is not None:
            ach = st["bytes"] / (st["ms"] * 1e-3) / 1e9
            roof["fused_step"] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                                  "avg_launch_ms": st["ms"] / st["launches"], "bytes_per_launch": st["bytes"] / st["launches"]}
        for c in classes.values():
            c["share"] = c["ms"] / total_ms

    # ---- final gather (the only collective of the path; outside the timed loop) -------------------
    gather_ms = None
    if world > 1:
        from composable_diffusion_models_b200 import dist as D
        # the collective itself: same-size warm-up (communicator / transport set-up, allocator) into a preallocated output
        buf = torch.empty((world * B,) + tuple(x.shape[1:]), device=dev)
        D.gather_samples(x, total=world * B, dst=None, out=buf)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = D.gather_samples(x, total=world * B, dst=None, out=buf)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = max_over_ranks(g0.elapsed_time(g1))
        assert full.shape[0] == world * B

    # ---- extra legs: N = 1 only (they would desynchronise the ranks of a scaling run) ---------------
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        def leg(name, fn):
            try:
                extras[name] = fn()
            except Exception as e:
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            torch.cuda.empty_cache()

        def full_chain():
            xs = torch.randn(B, *IMG, generator=gen).to(dev)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            xs = _sample_sde_chain(experts, weights, xs, CHAIN_STEPS, 1.0, "kernel", 7)
            c1.record()
            torch.cuda.synchronize()
            s = c0.elapsed_time(c1) * 1e-3
            return {"seconds": s, "samples_per_s": B / s, "steps": CHAIN_STEPS, "batch": B, "finite": bool(torch.isfinite(xs).all())}

        def other_modes():
            res = {}
            for prec in ("f16x3", "fp32"):
                if prec == args.precision:
                    continue
                try:
                    ex = make_experts(prec)
                    n = 5 if prec == "f16x3" else 2
                    ms_m, _, _ = device_resident(ex, B, n, 1)
                    res[prec] = {"ms_per_step": ms_m, "value": B / (CHAIN_STEPS * ms_m * 1e-3), "unit": "samples/s",
                                 "tflops": K_EXPERTS * B * UNET_GFLOP / ms_m}
                    del ex
                except Exception as e:
                    res[prec] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
            return res

        def grouped_vs_back_to_back():
            """K = 2 expert forwards with one grouped launch per convolution (cdm_unet_forward_grouped) against the same
            two forwards back to back, at the per-GPU batches of the strong-scaling reading (4096 / 8 = 512) and below."""
            from composable_diffusion_models_b200.models import forward_grouped
            lib = _lib.lib()
            res = {}
            for Bg in (128, 256, 512, 4096):
                xg = torch.randn(Bg, *IMG, device=dev)
                tg = torch.full((Bg,), 0.5, device=dev)
                row = {}
                for name, flag in (("grouped_ms", 1), ("back_to_back_ms", 0)):
                    lib.cdm_set_option(b"grouped", flag)
                    row[name] = _timed(lambda i: forward_grouped(experts, xg, tg), 20 if Bg <= 512 else 5, 3)
                lib.cdm_set_option(b"grouped", -1)
                row["speedup"] = row["back_to_back_ms"] / row["grouped_ms"]
                res[f"B{Bg}"] = row
            return res

        leg("grouped", grouped_vs_back_to_back)
        leg("full_chain", full_chain)
        leg("modes", other_modes)
        leg("gpu_eager_baseline", lambda: gpu_eager_baseline(dev, B))
        leg("parity", lambda: parity_block(dev, [args.precision] + [p for p in ("fp16", "f16x3", "fp32") if p != args.precision]))
        leg("configs", lambda: other_configs(dev, pk))

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            csteps = 4
            sec, kind = reference_cpu_step_seconds(CPU_BATCH, csteps, 1, threads)
            cpu = {"value": CPU_BATCH / (CHAIN_STEPS * sec), "unit": "samples/s", "cores": threads, "kind": kind,
                   "sample": f"{csteps} sampler steps at batch {CPU_BATCH} of the same workload, fp32, "
                             + ("the reference's own UNet + schedule modules (oracle/_ref)" if kind == "reference" else "oracle port")
                             + ", torch CPU ops"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "experts": K_EXPERTS, "batch_per_gpu": B,
                       "chain_steps": CHAIN_STEPS, "image": list(IMG), "noise": "in-kernel philox",
                       "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2",
                       "parallelism": f"dp{world} independent chains"},
            "tflops_per_step_algorithmic": K_EXPERTS * B * UNET_GFLOP / 1e3,
            "roofline": roof, "kernel_classes": classes, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": step_bytes, "d2h_bytes_per_step": step_bytes,
                    "ms_per_step": ms2 / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "final_gather_ms": gather_ms,
            "state_absmax_after_timed_steps": state_absmax, "strong": strong,
        }
        if "full_chain" in extras:
            line["full_chain_s"] = extras["full_chain"].get("seconds")
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
