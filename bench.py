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
over ranks.  See DESIGN.md "Measurement".
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


def _traffic():
    """DRAM bytes per conv launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the 10 tensor-core conv
    launches of one expert forward at B=4096) from the committed `ncu --set full` capture, or None."""
    p = os.path.join(ROOT, "profiles", "r01_conv_dram_traffic.json")
    try:
        return json.load(open(p))["avg_bytes_per_launch"]
    except Exception:
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


def oracle_step_seconds(batch, steps, threads):
    """The CPU restatement of the reference path (oracle/, kind 'port'): K UNet forwards + SDE update."""
    import torch
    from oracle import experts as E
    from oracle import samplers as OS
    torch.set_num_threads(threads)
    sds = [E.synth_state_dict(E.unet_small_spec(1), 1234 + k) for k in range(K_EXPERTS)]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, *IMG, generator=g)
    dt = 1.0 / CHAIN_STEPS

    def one(i, x):
        t = torch.full((batch,), 1.0 - i * dt)
        eps = [E.unet_small_forward(sd, x, t) for sd in sds]
        return OS.sde_step(x, eps, [1.0] * K_EXPERTS, 1.0 - i * dt, dt, 1.0, torch.randn(x.shape, generator=g))

    with torch.no_grad():
        x = one(0, x)
        t0 = time.perf_counter()
        for i in range(steps):
            x = one(1 + i, x)
        return (time.perf_counter() - t0) / steps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (oracle port;
    /root/reference itself does not exist on the GPU box)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    bcpu = 32
    steps = max(1, min(args.steps, 8))
    sec = oracle_step_seconds(bcpu, steps, threads)
    value = bcpu / (CHAIN_STEPS * sec)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mnist_unet_K2_sde_1000steps_B4096", "batch_timed": bcpu, "chain_steps": CHAIN_STEPS},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} sampler steps at batch {bcpu} (of the 1000-step, batch-4096 workload), fp32, torch CPU ops"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="samples per GPU (default: the named workload's 4096)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from composable_diffusion_models_b200 import _lib, steps as S
    from composable_diffusion_models_b200.compose_scores import sde_coefficients
    from composable_diffusion_models_b200.models import UNet

    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().cdm_device_check(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # synthetic experts: the reference architecture, default init under fixed seeds
    experts = []
    for k in range(K_EXPERTS):
        torch.manual_seed(1234 + k)
        experts.append(UNet(precision=args.precision).to(dev).eval())
    B = args.batch
    gen = torch.Generator(device="cpu").manual_seed(rank)
    x = torch.randn(B, *IMG, generator=gen).to(dev)
    dt = 1.0 / CHAIN_STEPS
    weights = [1.0] * K_EXPERTS

    from composable_diffusion_models_b200.compose_scores import _sample_sde_chain

    def run_steps(i0, n, x, z=None):
        """steps i0 .. i0+n-1 of the 1000-step chain through the public sampler path (cdm_unet_sample_sde: K forwards +
        the fused step per timestep, one host call); z = that step's injected noise, else in-kernel Philox (rank, i)."""
        i0 %= CHAIN_STEPS
        n = min(n, CHAIN_STEPS - i0)
        noise = "kernel" if z is None else (lambda i: z)
        return _sample_sde_chain(experts, weights, x, CHAIN_STEPS, 1.0, noise, rank, step_range=(i0, i0 + n))

    def step(i, x, z=None):
        return run_steps(i, 1, x, z)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) -----------------------------------------------------
    for i in range(warmup):
        x = step(i, x)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    e0.record()
    x = run_steps(warmup, args.steps, x)          # the K timed steps: one host call, kernels back to back
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    ms_per_step = ms / args.steps
    value = world * B / (CHAIN_STEPS * ms_per_step * 1e-3)
    state_absmax = float(x.abs().max())       # the chain state must stay finite (random-init experts make it grow)
    if not (state_absmax < float("inf")):
        raise RuntimeError("bench: the sampler state is not finite")

    # ---- end-to-end through the public API with host buffers (`e2e`) ------------------------------
    # every step: H2D of that step's injected noise from pinned memory, D2H of the step's result
    z_host = torch.randn(B, *IMG, generator=gen).pin_memory()
    x_host = torch.empty(B, *IMG).pin_memory()
    z_dev = torch.empty(B, *IMG, device=dev)
    for i in range(2):
        z_dev.copy_(z_host, non_blocking=True)
        x = step(i, x, z_dev)
        x_host.copy_(x, non_blocking=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        z_dev.copy_(z_host, non_blocking=True)
        x = step(i, x, z_dev)
        x_host.copy_(x, non_blocking=True)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms2], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms2 = float(tm.item())
    e2e_value = world * B / (CHAIN_STEPS * (ms2 / args.steps) * 1e-3)
    step_bytes = B * IMG[0] * IMG[1] * IMG[2] * 4

    # ---- per-kernel-class timing (roofline) on rank 0: a separate short pass, events around every launch
    roof, classes = None, None
    if rank == 0:
        nprof = min(args.steps, 3)
        _lib.prof_enable(True)
        for i in range(nprof):
            x = step(i, x)
        torch.cuda.synchronize()
        classes = _lib.prof_summary()
        _lib.prof_enable(False)
        pk = _peaks()
        total_ms = sum(c["ms"] for c in classes.values())
        conv = classes.get("conv_tc") or classes.get("conv_fp32")
        if conv:
            ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
            roof = {"kernel": ("conv_halo_kernel / conv_stack3_kernel / conv_tc_kernel (tcgen05 implicit-GEMM 3x3 convs, all launches)"
                               if "conv_tc" in classes else "conv_fp32_kernel"),
                    "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                    "traffic": _traffic() if "conv_tc" in classes and B == BATCH else None,
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
        D.gather_samples(x[:8].contiguous(), total=world * 8, dst=0)      # communicator / transport set-up is not the gather
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = D.gather_samples(x, total=world * B, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        assert (full is None) == (rank != 0)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            bcpu, csteps = 32, 6
            sec = oracle_step_seconds(bcpu, csteps, threads)
            cpu = {"value": bcpu / (CHAIN_STEPS * sec), "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": f"{csteps} sampler steps at batch {bcpu} of the same workload, fp32 oracle (torch CPU ops)"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "mnist_unet_K2_sde_1000steps_B4096", "experts": K_EXPERTS, "batch_per_gpu": B,
                       "chain_steps": CHAIN_STEPS, "image": list(IMG), "noise": "in-kernel philox",
                       "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2",
                       "parallelism": f"dp{world} independent chains"},
            "tflops_per_step_algorithmic": K_EXPERTS * B * UNET_GFLOP / 1e3,
            "roofline": roof, "kernel_classes": classes, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": step_bytes, "d2h_bytes_per_step": step_bytes,
                    "ms_per_step": ms2 / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "final_gather_ms": gather_ms,
            "state_absmax_after_timed_steps": state_absmax,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
