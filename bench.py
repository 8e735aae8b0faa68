#!/usr/bin/env python
"""bench.py — mel-spectrograms/sec of the DDPM sampling hot path (BASELINE.json metric).

A "step" is one denoising step (U-Net forward + fused DDPM update) over one batch of synthetic input:
config C2 of BASELINE.json — audio-diffusion-256 architecture (scripts/train_unet.py:115-137), 256x256x1,
batch 64 per GPU, DDPM with 1000 steps per mel-spectrogram.  value = (batch * n_gpus) / (1000 * step time).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (libb200ad.so)
    python bench.py --impl reference --gpus N ...            # CPU arm: oracle port of diffusers, host cores

Weights are random-init (no network for checkpoints), inputs synthetic (seed 42, as train_unet.py:314).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

REF_ARCH = dict(
    in_channels=1, out_channels=1, layers_per_block=2, block_out_channels=(128, 128, 256, 256, 512, 512),
    down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"))
GFLOP_PER_SAMPLE_FWD = 496.415145984   # analytic, oracle.unet_oracle.unet_flops(256, 256); SURVEY §8(d)
DDPM_STEPS = 1000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1439.1), d.get("hbm_gbs", 6572.5), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_rate(batch: int, hw: int, steps: int, warmup: int):
    """Oracle (port of diffusers UNet2DModel + DDPMScheduler.step) on the host cores; returns mel-spec/s."""
    from oracle.schedulers_oracle import OracleDDPM
    from oracle.unet_oracle import UNetConfig, init_weights, unet_forward
    cfg = UNetConfig(sample_size=(hw, hw))
    w = init_weights(cfg, seed=0)
    sch = OracleDDPM()
    sch.set_timesteps(DDPM_STEPS)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(batch, 1, hw, hw, generator=g)
    ts = sch.timesteps
    with torch.no_grad():
        for i in range(warmup):
            eps = unet_forward(w, cfg, x, ts[i])
            x = sch.step(eps, ts[i], x, generator=g)["prev_sample"]
        t0 = time.perf_counter()
        for i in range(steps):
            eps = unet_forward(w, cfg, x, ts[warmup + i])
            x = sch.step(eps, ts[warmup + i], x, generator=g)["prev_sample"]
        dt = (time.perf_counter() - t0) / steps
    return batch / (DDPM_STEPS * dt), dt


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    if torch.get_num_threads() == 1 and (os.cpu_count() or 1) > 2:
        torch.set_num_threads(max(1, os.cpu_count() // 2))  # torchrun forces OMP_NUM_THREADS=1: use the physical cores
    cores = torch.get_num_threads()
    batch, hw = 4, 256
    steps = max(1, min(args.steps, 3))
    warmup = 1
    val, dt = cpu_step_rate(batch, hw, steps, warmup)
    line = {
        "impl": "reference", "metric": "mel-spectrograms/sec (1000-step DDPM, 256x256x1)", "value": val,
        "unit": "mel-spectrograms/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (random-init weights, seed-42 noise)",
        "config": {"workload": "audio-diffusion-256 DDPM 1000-step, 256x256x1 (C2 architecture)",
                   "note": "CPU arm: bounded sample"},
        "cpu_baseline": {"value": val, "unit": "mel-spectrograms/s", "cores": cores, "kind": "port",
                         "sample": f"batch {batch} x {steps} denoise steps of 1000 at 256x256, oracle port of "
                                   f"diffusers UNet2DModel+DDPMScheduler (diffusers/librosa not installable), "
                                   f"{cores} torch threads of {os.cpu_count()} cpus"},
        "e2e": {"value": val, "unit": "mel-spectrograms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    rank, world, local = dist_env()
    from audio_diffusion_b200 import _lib
    from audio_diffusion_b200.schedulers import DDPMScheduler
    from audio_diffusion_b200.unet import UNet2DModel
    _lib.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, HW = args.batch, args.res
    torch.manual_seed(0)
    model = UNet2DModel(sample_size=(HW, HW), seed=0, **REF_ARCH).to(dev)
    if use_dist:
        # batched sampling shards over GPUs: ONE broadcast of the weights at init, no per-step collective
        from audio_diffusion_b200.parallel import broadcast_parameters
        broadcast_parameters(model.parameters(), src=0)
    sch = DDPMScheduler()
    sch.set_timesteps(DDPM_STEPS)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    x = torch.randn(B, 1, HW, HW, generator=g, device=dev)
    z = torch.randn(B, 1, HW, HW, generator=g, device=dev)
    ts = sch.timesteps
    K, W = args.steps, max(args.warmup, 3)

    def step(i, xin, xout):
        t = ts[i % DDPM_STEPS]
        return model.forward_step(xin, t, sch.step_coef(t), noise=z if int(t) > 0 else None, out=xout)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(W):
            step(i, x, x)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(W + i, x, x)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / K
        clocks = sampler.stop()
        launches = model.last_launch_count * K

        # ---- e2e: the same step through the public API with HOST (pinned) buffers, copies inside the timed region
        xh = torch.randn(B, 1, HW, HW).pin_memory()
        zh = torch.randn(B, 1, HW, HW).pin_memory()
        oh = torch.empty(B, 1, HW, HW).pin_memory()
        Ke = max(3, min(K, 10))
        for rep in range(2):
            barrier()
            e0.record()
            for i in range(Ke):
                xd = xh.to(dev, non_blocking=True)
                zd = zh.to(dev, non_blocking=True)
                t = ts[(W + i) % DDPM_STEPS]
                out = model.forward_step(xd, t, sch.step_coef(t), noise=zd)
                oh.copy_(out, non_blocking=True)
            e1.record()
            barrier()
        ms_e2e = e0.elapsed_time(e1) / Ke

        # ---- roofline of the dominant kernel (conv_tc_kernel), CUDA events around every launch of one step
        prof = None
        if rank == 0:
            L = _lib.lib()
            maxops = 1024
            op_ms = (C.c_float * maxops)()
            op_kind = (C.c_int * maxops)()
            op_fl = (C.c_double * maxops)()
            t = ts[5]
            tt = model._timesteps(t, B, dev)
            coef = sch.step_coef(t)
            tot = {}
            for rep in range(3):
                n = L.b200ad_unet_profile_step(model._h, x.data_ptr(), tt.data_ptr(), z.data_ptr(), C.byref(coef),
                                               x.data_ptr(), op_ms, op_kind, op_fl, maxops, _lib.stream_ptr())
                if n < 0:
                    raise RuntimeError(L.b200ad_last_error().decode())
                if rep == 0:
                    continue
                for i in range(n):
                    k = op_kind[i]
                    a = tot.setdefault(k, [0.0, 0.0, 0])
                    a[0] += op_ms[i] / 2
                    a[1] += op_fl[i] / 2
                    a[2] += 0.5
            prof = {"ops": n, "by_kind": tot,
                    "per_op": [(int(op_kind[i]), float(op_ms[i]), float(op_fl[i])) for i in range(n)]}

    ms_t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if use_dist:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = ms_t.tolist()
    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    peak_tf, peak_hbm, how = peaks()
    value = B * world / (DDPM_STEPS * ms * 1e-3)
    e2e_val = B * world / (DDPM_STEPS * ms_e2e * 1e-3)
    names = {0: "temb", 1: "conv_in", 2: "gn_finalize", 3: "conv_tc", 4: "upsample", 5: "parity_split", 6: "attention", 7: "conv_out"}
    conv = prof["by_kind"].get(3, [0.0, 0.0, 0])
    step_prof_ms = sum(v[0] for v in prof["by_kind"].values())
    conv_tf = conv[1] / (conv[0] * 1e-3) / 1e12 if conv[0] > 0 else 0.0
    # DRAM bytes per conv_tc launch come from the committed ncu launch list of this same command (profiles/); they are
    # only valid for the default workload (batch 64, 256x256) and are null otherwise
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r01_final.json")
    if os.path.exists(tpath) and HW == 256 and B == 64:
        with open(tpath) as f:
            tk = json.load(f)["kernels"].get("conv_tc_kernel")
        if tk:
            traffic = tk["dram_read_bytes_per_launch"] + tk["dram_write_bytes_per_launch"]
            traffic_src = "profiles/traffic_r01_final.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over 203 launches)"
    roof = {"bound": "tensor", "kernel": "conv_tc_kernel", "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": conv_tf / peak_tf, "peak_source": f"{how} bf16 sustained (MEASURED_PEAKS.json)",
            "traffic": traffic, "traffic_source": traffic_src,
            "launches_per_step": int(conv[2]), "avg_launch_ms": conv[0] / max(conv[2], 1),
            "algorithmic_gflop_per_launch": conv[1] / max(conv[2], 1) / 1e9,
            "kernel_share_of_step": conv[0] / step_prof_ms if step_prof_ms else None,
            "whole_step_tflops": GFLOP_PER_SAMPLE_FWD * B / (ms * 1e-3) / 1e3 if HW == 256 else None,
            "whole_step_frac": (GFLOP_PER_SAMPLE_FWD * B / (ms * 1e-3) / 1e3 / peak_tf) if HW == 256 else None,
            "ms_by_kernel": {names[k]: round(v[0], 4) for k, v in sorted(prof["by_kind"].items())}}
    cpu = None
    if not args.no_cpu and world == 1:
        cores = torch.get_num_threads()
        v, dt = cpu_step_rate(4, HW, 3, 1)
        cpu = {"value": v, "unit": "mel-spectrograms/s", "cores": cores, "kind": "port",
               "sample": f"batch 4 x 3 denoise steps (of 1000) at {HW}x{HW} on {cores} torch threads "
                         f"({os.cpu_count()} cpus); oracle port of diffusers UNet2DModel + DDPMScheduler.step"}
    line = {
        "metric": "mel-spectrograms/sec (1000-step DDPM, 256x256x1)", "value": value, "unit": "mel-spectrograms/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 operands, fp32 accumulate (fp32 master weights, fp64 GroupNorm sums)",
        "data": "synthetic (random-init weights seed 0, noise seed 42)",
        "config": {"workload": f"audio-diffusion-256 DDPM 1000-step, batch={B} per GPU, {HW}x{HW}x1, {world}xB200",
                   "global_batch": B * world, "parallelism": f"dp{world} (batch sharded, weights broadcast once)",
                   "l2": "activations per step (>10 GB) far exceed the 126 MB L2; no explicit flush",
                   "step": "one denoise step = UNet2DModel forward + fused DDPMScheduler.step"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_val, "unit": "mel-spectrograms/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * B * HW * HW * 4, "d2h_bytes_per_step": B * HW * HW * 4,
                "api": "UNet2DModel.forward_step on pinned host tensors (x, z in; x_prev out)"},
        "roofline": roof, "cpu_baseline": cpu,
    }
    if args.dump_ops:
        os.makedirs(os.path.dirname(args.dump_ops) or ".", exist_ok=True)
        json.dump(prof["per_op"], open(args.dump_ops, "w"))
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ training mode (extra)
def run_train(args):
    """`--mode train`: one iteration of scripts/train_unet.py:238-267 per step (config C5 of SURVEY §8: 256x256, bf16
    activations, data parallel). Not the headline metric — a separate JSON line with its own metric name."""
    rank, world, local = dist_env()
    from audio_diffusion_b200 import _lib
    from audio_diffusion_b200.parallel import broadcast_parameters
    from audio_diffusion_b200.schedulers import DDPMScheduler
    from audio_diffusion_b200.training import EMAModel, FusedAdamW, train_step
    from audio_diffusion_b200.unet import UNet2DModel
    _lib.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch if args.batch != 64 else 16          # per-GPU batch (C5 states it): default 16
    HW = args.res
    model = UNet2DModel(sample_size=(HW, HW), seed=0, **REF_ARCH).to(dev).train()
    if use_dist:
        broadcast_parameters(model.parameters(), src=0)
    opt = FusedAdamW(model.parameters(), lr=1e-4, betas=(0.95, 0.999), weight_decay=1e-6, eps=1e-8, max_grad_norm=1.0)
    ema = EMAModel(model.parameters(), inv_gamma=1.0, power=0.75, max_value=0.9999)
    opt.attach_ema(ema)
    sch = DDPMScheduler()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # synthetic dataset batch: uint8 images through ToTensor + Normalize(.5, .5) (train_unet.py:73-78), resident on the GPU
    x = (torch.randint(0, 256, (B, 1, HW, HW), device=dev, generator=gen).float() / 255.0 - 0.5) / 0.5
    W = max(args.warmup, 3)
    for _ in range(W):
        train_step(model, opt, sch, x, ema=ema, generator=gen)
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = train_step(model, opt, sch, x, ema=ema, generator=gen)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    ms_t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if use_dist:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.barrier()
    ms = ms_t.item()
    if rank == 0:
        peak_tf, _, how = peaks()
        gf = GFLOP_PER_SAMPLE_FWD if HW == 256 else None
        tf = 3 * gf * B / ms if gf else None
        print(json.dumps({
            "metric": "images/sec (train_unet.py step: fwd + bwd + clip + AdamW + EMA, 256x256x1)", "mode": "train",
            "value": B * world / ms * 1e3, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 activations and activation gradients, fp32 parameters / parameter gradients / optimizer",
            "data": "synthetic (uint8 images -> [-1,1], random-init weights seed 0)",
            "config": {"workload": f"train_unet.py iteration, batch={B} per GPU, {HW}x{HW}x1, {world}xB200",
                       "global_batch": B * world, "parallelism": f"dp{world} (one all-reduce of the flat gradient buffer per step)"},
            "clocks": clocks, "gpu_launches": (model.last_launch_count + model.last_backward_launch_count + 4) * args.steps,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": tf / peak_tf if tf else None, "peak_source": f"{how} bf16 sustained (MEASURED_PEAKS.json)",
                         "traffic": None, "flops_model": "3 x 496.42 GFLOP per sample (SURVEY §8d)"},
            "loss": float(loss)}))
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="sample", choices=["sample", "train"],
                    help="sample = the headline metric (default); train = one train_unet.py iteration per step")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch profile (kind, ms, flops) to this JSON")
    args = ap.parse_args()
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version banner to stdout)
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "train":
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
