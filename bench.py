#!/usr/bin/env python
"""
bench.py — denoiser frame-steps/s (sampling) of the FDM hot path on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]

A "step" is one diffusion step of the sampler over one batch: UNetVideoModel forward (the whole kernel schedule)
+ the fused posterior update + the step's Gaussian noise draw, i.e. what `diffusion.p_sample` does.  One step
processes B*K frame-steps per GPU.  Multi-GPU = the video batch sharded over ranks, no data-path collective
(weak scaling: B videos per GPU fixed), launched by torchrun (one process per GPU).

Lines printed (rank 0, ONE JSON line): see README/DESIGN.md §measurement for every key.
  value     : whole-job frame-steps/s with inputs resident in HBM (CUDA events, L2 flushed between steps, max over ranks)
  e2e       : same metric through diffusion.p_sample_loop with pinned HOST inputs, H2D + D2H inside the timed region
  roofline  : the dominant kernel class (implicit-GEMM convs) timed alone with CUDA events: algorithmic FLOPs / time
  cpu_baseline : the CPU oracle (port of the reference path) timed on this box's host cores on a bounded sample
  --impl reference : times that CPU path only (rank 0), same workload/metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch as th  # noqa: E402

PIXEL = dict(diffusion_space="pixel", pre_encoded=False, pre_encoded_stats_dict=None)

# BASELINE.json configs.  `B` is videos PER GPU (weak scaling).
WORKLOADS = {
    # cfg4: hierarchy-2 sampling, 300-frame video, batch 64 sharded over 8 GPUs -> 8 videos/GPU, stage of K=20 frames,
    # Carla-latent model of cfg2 (nc=64, nrb=1, 4x32x32 latents), 1000-step schedule.  The sampling headline.
    "cfg4-sampling": dict(over=dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000),
                          B=8, K=20, n_obs=10, video_len=300),
    # cfg2 shape (batch 1, K=5): launch-latency bound by construction
    "cfg2-sampling-b1": dict(over=dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000),
                             B=1, K=5, n_obs=3, video_len=20),
    # cfg1: the reference's own CPU-runnable case
    "cfg1-sampling": dict(over=dict(image_size=32, in_channels=4, num_channels=32, num_res_blocks=1, diffusion_steps=32),
                          B=1, K=5, n_obs=3, video_len=20),
    # cfg5: long-context temporal attention stress (nc=128, K=40, 4x64x64 latents)
    "cfg5-sampling": dict(over=dict(image_size=64, in_channels=4, num_channels=128, num_res_blocks=1, diffusion_steps=1000),
                          B=2, K=40, n_obs=20, video_len=300),
    # cfg3's model used as a sampler (pixel space, 128x128x3, nc=128, K=20, 2 videos per GPU): the 128-wide halo convs
    "cfg3-sampling": dict(over=dict(image_size=128, in_channels=3, num_channels=128, num_res_blocks=1, diffusion_steps=1000),
                          B=2, K=20, n_obs=10, video_len=300),
}
DEFAULT_WORKLOAD = "cfg4-sampling"


def random_state_dict(model, seed=1):
    """Random NON-ZERO weights (a freshly constructed model outputs eps == 0: zero_module, reference nn.py:68-74)."""
    g = th.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if v.dim() > 1:
            sd[k] = th.randn(v.shape, generator=g) / (v[0].numel() ** 0.5)
        elif k.endswith(".weight"):
            sd[k] = 1.0 + 0.1 * th.randn(v.shape, generator=g)
        else:
            sd[k] = 0.1 * th.randn(v.shape, generator=g)
    return sd


def synthetic_batch(cfg_over, B, K, n_obs, video_len, seed):
    g = th.Generator().manual_seed(seed)
    C, S = cfg_over["in_channels"], cfg_over["image_size"]
    x0 = th.randn(B, K, C, S, S, generator=g).clamp(-1, 1)
    fi = th.stack([th.sort(th.randperm(video_len, generator=g)[:K]).values for _ in range(B)]).long()
    obs = th.zeros(B, K, 1, 1, 1)
    obs[:, :n_obs] = 1
    return dict(x0=x0, frame_indices=fi, obs_mask=obs, latent_mask=1 - obs)


def build_native(over, device):
    from improved_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults
    d = model_and_diffusion_defaults()
    d.update(over)
    d["diffusion_space_kwargs"] = dict(PIXEL)
    model, diffusion = create_model_and_diffusion(**d)
    sd = random_state_dict(model)
    model.load_state_dict(sd, strict=True)
    model.to(device).eval()
    return model, diffusion, sd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [s.strip() for s in line.split(",")]
                if len(c) < 7:
                    continue
                try:
                    sm.append(float(c[0]))
                    mx.append(float(c[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_port_steps(over, B, K, batch, sd, n_steps, warmup, threads):
    """The reference path on host cores: the oracle port (oracle/fdm_oracle.py), fp32 torch CPU ops.  Checker/baseline only."""
    from oracle import fdm_oracle as O
    th.set_num_threads(threads)
    cfg = O.make_cfg(**over)
    tab = O.Tables(cfg)
    g = th.Generator().manual_seed(123)
    x = th.randn(batch["x0"].shape, generator=g)
    times = []
    n = tab.num_timesteps
    with th.no_grad():
        for s in range(warmup + n_steps):
            t = th.full((B,), n - 1 - (s % n), dtype=th.int64)
            noise = th.randn(x.shape, generator=g)
            t0 = time.perf_counter()
            out = O.p_sample(tab, sd, cfg, x, t, noise, batch)
            dt = time.perf_counter() - t0
            x = out["sample"]
            if s >= warmup:
                times.append(dt)
    return times


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; /root/reference cannot travel to
    the GPU box), all host threads, same workload / metric.  Rank 0 only."""
    if rank != 0:
        return
    over, B, K = wl["over"], wl["B"], wl["K"]
    from improved_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults
    d = model_and_diffusion_defaults()
    d.update(over)
    d["diffusion_space_kwargs"] = dict(PIXEL)
    model, _ = create_model_and_diffusion(**d)
    sd = random_state_dict(model)
    batch = synthetic_batch(over, B, K, wl["n_obs"], wl["video_len"], seed=0)
    threads = os.cpu_count() or 1
    times = cpu_port_steps(over, B, K, batch, sd, args.steps, args.warmup, threads)
    total = sum(times)
    v = B * K * len(times) / total
    line = {"impl": "reference", "metric": "denoiser frame-steps/sec (sampling)", "value": v, "unit": "frame-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, **{k: wl[k] for k in ("B", "K", "n_obs")}, **over},
            "cpu_baseline": {"value": v, "unit": "frame-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{len(times)} diffusion steps of the same B={B},K={K} batch (oracle port, fp32, "
                                       f"torch CPU ops, {threads} threads)"},
            "e2e": {"value": v, "unit": "frame-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _peak_tflops():
    """(dense bf16 TFLOP/s, source): the driver-measured cuBLAS number of this pool's B200s, else the profiling recipe's fallback"""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]), "measured"
    except Exception:
        return 1590.0, "fallback"


TRAIN_WORKLOADS = {
    # BASELINE.json configs[1]: Carla-latent, nc = 64, K = 5, batch 1, on 1 B200
    "cfg2-train": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000), 1, 5, 20),
    # BASELINE.json configs[2]: Carla pixel-scale, nc = 128, K = 20, 128x128x3 frames, batch 2 PER GPU (DDP over the box)
    "cfg3-train": (dict(image_size=128, in_channels=3, num_channels=128, num_res_blocks=1, diffusion_steps=1000), 2, 20, 6),
    # BASELINE.json configs[4] as a training shape (not part of the default run): long-context model, K = 40, 4x64x64 latents
    "cfg5-train": (dict(image_size=64, in_channels=4, num_channels=128, num_res_blocks=1, diffusion_steps=1000), 2, 40, 6),
    # cfg4's shard shape as a training step (8 videos x 20 frames of 4x32x32 latents, nc = 64)
    "cfg4-train": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000), 8, 20, 10),
}


def bench_train(dev, world, precision, workload="cfg2-train", warmup=3):
    """train samples/s: one training_losses forward + backward (+ DDP gradient allreduce over NCCL when world > 1) + AdamW step
    per iteration, per-GPU batch fixed (weak scaling).  Forward AND backward run on the native kernel schedules
    (engine._DenoiserFn, CUDA-graph replays); the same step through torch.autograd over the PyTorch expression of the network
    (cuDNN/cuBLAS under autocast, FDM_TRAIN_ENGINE=autograd) is timed beside it as `autograd_ms_per_step`."""
    import torch.distributed as dist
    over, B, K, steps = TRAIN_WORKLOADS[workload]
    model, diffusion, _ = build_native(over, dev)
    model.precision = precision
    model.train()
    net = model
    if world > 1:
        from improved_diffusion import sharding
        # FDM_DDP=torch: torch DistributedDataParallel exactly as train_util.py:118-125 (drop-in path, per-parameter hooks);
        # default: one allreduce over the native backward's flat gradient buffer
        net = sharding.wrap_ddp(model, dev) if os.environ.get("FDM_DDP", "flat") == "torch" else sharding.FlatGradDataParallel(model)
    # FDM_OPT=torch: torch.optim.AdamW(fused=True) (multi-tensor kernels over the 390 tensors); default: optim.FlatAdamW — the same
    # update rule as ONE fdm_adamw launch over flat parameter / moment / gradient buffers
    if os.environ.get("FDM_OPT", "flat") == "torch":
        opt = th.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.0, fused=True)
    else:
        from improved_diffusion.optim import FlatAdamW
        # flat-gradient mode (model=...) needs the flat data-parallel wrapper, not torch DDP's per-parameter hooks
        opt = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=0.0,
                        model=None if (world > 1 and os.environ.get("FDM_DDP", "flat") == "torch") else model)
    batch = {k: v.to(dev) for k, v in synthetic_batch(over, B, K, 3, 4 * K, seed=1 + int(os.environ.get("RANK", "0"))).items()}
    g = th.Generator(device=dev).manual_seed(0)

    def step():
        t = th.randint(0, diffusion.num_timesteps, (B,), device=dev, generator=g)
        terms = diffusion.training_losses(net, batch["x0"], t, model_kwargs=batch, latent_mask=1 - batch["obs_mask"],
                                          eval_mask=batch["latent_mask"])
        opt.zero_grad(set_to_none=True)
        terms["loss"].mean().backward()
        opt.step()
        return terms["loss"]

    def timed(engine, n):
        os.environ["FDM_TRAIN_ENGINE"] = engine
        for _ in range(warmup):
            step()
        th.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = step()
        e1.record()
        th.cuda.synchronize()
        assert bool(th.isfinite(loss).all()), "training loss is not finite"
        ms = th.tensor([e0.elapsed_time(e1)], device=dev, dtype=th.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    ms_auto = timed("autograd", max(2, steps // 2)) if world == 1 else None  # the A/B arm is a single-GPU comparison
    th.cuda.empty_cache()
    ms = timed("native", steps)
    plan = next(iter(model.engine().train_plans.values()))
    flops = plan.flops + plan.bflops
    out = {"metric": "train samples/sec", "value": world * B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
           "autograd_ms_per_step": ms_auto, "steps": steps,
           "config": {"workload": workload, "batch_per_gpu": B, "frames": K, **over, "optimizer": "AdamW, weight_decay 0 (" + type(opt).__name__ + ")"},
           "launches_per_step": {"forward": plan.n_launches, "backward": plan.n_bwd_launches},
           "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "peak": _peak_tflops()[0],
                        "peak_source": _peak_tflops()[1], "frac": flops / (ms * 1e-3) / 1e12 / _peak_tflops()[0],
                        "flops_per_step": flops, "what": "algorithmic fwd + dgrad + wgrad + attention FLOPs of one step / step time "
                                                         "(optimizer, loss and host work included in the time)"},
           "path": "native: forward + backward kernel schedules of libfdm_sm100.so behind one autograd node (conv dgrad and wgrad on "
                   "tcgen05, GroupNorm / attention / RPENet backward kernels); n_gpus > 1: ONE NCCL allreduce over the flat gradient "
                   "buffer per step (sharding.FlatGradDataParallel; FDM_DDP=torch selects torch DDP as in train_util.py:118-125)"}
    del opt, net
    if world == 1:
        try:
            out["e2e"] = _train_e2e(model, diffusion, over, B, K, steps)
        except Exception as ex:  # the device-resident number above stands on its own; say what went wrong instead of hiding it
            out["e2e"] = {"error": repr(ex)[:300]}
    del model
    th.cuda.empty_cache()
    return out


def _train_e2e(model, diffusion, over, B, K, steps):
    """The same training step end to end through train_step.NativeTrainStep with HOST video batches (3K frames each): per step
    mask sampling (the reference's draw order), frame gather, pinned upload, native forward + backward, gradient norm, FlatAdamW +
    EMA, and one device->host read of the logged scalars (left in flight: `defer=True`).  Wall clock around the loop."""
    import numpy as np
    from improved_diffusion.train_step import NativeTrainStep
    os.environ["FDM_TRAIN_ENGINE"] = "native"
    runner = NativeTrainStep(model, diffusion, lr=1e-4, max_frames=K, ema_rate="0.9999")
    C, S, T = over["in_channels"], over["image_size"], 3 * K
    g = th.Generator().manual_seed(5)
    pool = [th.randn(B, T, C, S, S, generator=g).clamp(-1, 1) for _ in range(4)]
    th.manual_seed(6)
    np.random.seed(6)
    for i in range(3):
        runner.run_step(pool[i % 4], pool[(i + 1) % 4], defer=True)
    runner.flush()
    th.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        runner.run_step(pool[i % 4], pool[(i + 1) % 4], defer=True)
    rec = runner.flush()
    th.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    assert np.isfinite(rec["loss"]) and np.isfinite(rec["grad_norm"]), rec
    return {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
            "h2d_bytes_per_step": B * K * (C * S * S * 4 + 8 + 4 + 4), "d2h_bytes_per_step": (4 * B + 1) * 4,
            "what": f"NativeTrainStep.run_step on host batches [B={B}, T={T}, {C}, {S}, {S}] fp32 (masks, gather, upload, fwd, bwd, "
                    "grad norm, AdamW+EMA, log read)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    if args.warmup < 3:
        args.warmup = 3
    import torch.distributed as dist
    if not th.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from improved_diffusion import _native as N_
    over, B, K = wl["over"], wl["B"], wl["K"]
    model, diffusion, sd = build_native(over, dev)
    model.precision = args.precision
    batch = synthetic_batch(over, B, K, wl["n_obs"], wl["video_len"], seed=rank)  # each rank its own shard of videos
    shape = tuple(batch["x0"].shape)
    C, S = over["in_channels"], over["image_size"]

    # ------------------------------------------------------------------ device-resident step (value)
    eng = model.engine()
    P = eng.plan_for(B, K, S, S, dev)
    kw_dev = {k: v.to(dev) for k, v in batch.items()}
    eng.load_conditioning(P, kw_dev["x0"], kw_dev["frame_indices"], kw_dev["obs_mask"], kw_dev["latent_mask"])
    tb = diffusion._tables(dev)
    P.set_t_source(tb["model_t"])
    nbuf = th.empty(shape, device=dev)
    step_args = N_.DdpmStepArgs(x=P.ptr(P.x), eps=P.ptr(P.eps), noise=nbuf.data_ptr(), coef=tb["step"].data_ptr(),
                                t=P.ptr(P.t_index), sample=P.ptr(P.x), pred_xstart=None, per_video=K * C * S * S, B=B, clip=1)
    stream = th.cuda.current_stream(dev)

    def body():
        s = th.cuda.current_stream(dev).cuda_stream
        P.run(s)
        N_.call("fdm_ddpm_step", step_args, s)

    P.x_view.normal_()
    P.t_index_view.fill_(diffusion.num_timesteps - 1)
    nbuf.normal_()
    body()
    stream.synchronize()
    graph = th.cuda.CUDAGraph()
    with th.cuda.graph(graph):
        body()
    flush = th.empty(192 * 1024 * 1024, dtype=th.uint8, device=dev)  # > 126 MB L2

    def one_step(i):
        P.t_index_view.fill_(i)
        nbuf.normal_()
        graph.replay()

    n_t = diffusion.num_timesteps
    P.x_view.normal_()
    for w in range(args.warmup):
        one_step(n_t - 1 - (w % n_t))
    th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    evs = [(th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        evs[k][0].record()
        one_step(n_t - 1 - ((args.warmup + k) % n_t))
        evs[k][1].record()
    th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    clk = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = th.tensor([sum(step_ms)], device=dev, dtype=th.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = world * B * K * args.steps / (total_ms * 1e-3)
    assert bool(th.isfinite(P.x_view).all()), "non-finite sampler state"

    # hot (no L2 flush, back-to-back) for information
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        one_step(n_t - 1 - (k % n_t))
    e1.record()
    th.cuda.synchronize()
    hot_ms = e0.elapsed_time(e1) / args.steps

    # ------------------------------------------------------------------ roofline of the dominant kernel class (convs), timed alone
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf, peak_src = (peaks["bf16_tflops"], "measured") if "bf16_tflops" in peaks else (1590.0, "fallback")
    import ctypes as C_
    sp = C_.c_void_p(stream.cuda_stream)

    def is_halo(st):
        """launches served by conv_halo_kernel (conv_halo.cu): 3x3 stride-1 tcgen05 convs on 16/32/64/128-wide maps"""
        if not (st.engine == N_.CONV_TC and st.ksize == 3 and st.stride == 1 and not st.upsample and not st.out_nchw):
            return False
        return st.Win in (16, 32, 64, 128) and st.Hin % max(1, 128 // st.Win) == 0 and st.Cout >= 32 and st.Cout % 4 == 0 and st.C0 % 8 == 0

    def conv_flops(st):
        pad = st.ksize // 2
        ho = (st.Hin + 2 * pad - st.ksize) // st.stride + 1
        wo = (st.Win + 2 * pad - st.ksize) // st.stride + 1
        return 2 * st.N * ho * wo * st.Cout * (st.ksize * st.ksize * st.C0 + st.C1)

    def time_calls(calls, reps=20):
        for fn, ref in calls:
            fn(ref, sp)
        th.cuda.synchronize()
        c0, c1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(reps):
            for fn, ref in calls:
                fn(ref, sp)
        c1.record()
        th.cuda.synchronize()
        return c0.elapsed_time(c1) / reps

    convs = [(fn, ref, st) for (name, fn, ref), st in zip(P.calls, P._structs) if name == "fdm_conv"]
    halo = [(fn, ref, st) for fn, ref, st in convs if is_halo(st)]
    conv_ms = time_calls([(fn, ref) for fn, ref, _ in convs])
    conv_tf = P.conv_flops / (conv_ms * 1e-3) / 1e12
    roofline = None
    if halo:
        # the dominant kernel: conv_halo_kernel.  achieved = algorithmic FLOPs per launch / average launch duration, measured
        # live with CUDA events over its launches of one step run back-to-back on the stream they are launched on
        halo_ms = time_calls([(fn, ref) for fn, ref, _ in halo])
        halo_fl = sum(conv_flops(st) for _, _, st in halo)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_halo_traffic.json")))
            if tj.get("workload") == args.workload:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        achieved = halo_fl / (halo_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_halo_kernel (3x3 stride-1 implicit-GEMM convs, tcgen05)", "achieved": achieved,
                    "peak": peak_tf, "peak_source": peak_src, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                    "launches_per_step": len(halo), "flops_per_launch": halo_fl / len(halo), "us_per_launch": 1e3 * halo_ms / len(halo),
                    "share_of_step": halo_ms / hot_ms, "share_of_step_flops": halo_fl / P.flops}
    all_convs = {"achieved": conv_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf, "launches_per_step": len(convs),
                 "ms_per_step": conv_ms, "share_of_step": conv_ms / hot_ms, "flops_per_step": P.conv_flops}
    step_tf = P.flops / (ms_per_step * 1e-3) / 1e12
    step_roofline = {"bound": "tensor", "achieved": step_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": step_tf / peak_tf,
                     "flops_per_frame_step": P.flops / (B * K)}

    # ------------------------------------------------------------------ e2e through the public API (host buffers)
    e2e = None
    if not args.no_e2e:
        host = {k: v.pin_memory() for k, v in batch.items()}
        out_host = th.empty(shape, dtype=th.float32).pin_memory()

        def stage():
            kw = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            final, _ = diffusion.p_sample_loop(model, shape, clip_denoised=True, model_kwargs=kw, latent_mask=kw["latent_mask"])
            out_host.copy_(final, non_blocking=True)
            th.cuda.synchronize()

        stage()  # warm-up (captures the sampler graph)
        th.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        stage()
        if world > 1:
            dist.barrier()
        dt = th.tensor([time.perf_counter() - t0], device=dev, dtype=th.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = out_host.numel() * 4
        e2e = {"value": world * B * K * n_t / dt, "unit": "frame-steps/s", "h2d_bytes_per_step": h2d / n_t,
               "d2h_bytes_per_step": d2h / n_t, "what": f"diffusion.p_sample_loop, one stage of {n_t} steps, pinned host "
               f"inputs -> device -> pinned host result, wall clock incl. copies", "seconds": dt}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        times = cpu_port_steps(over, B, K, batch, sd_cpu, n_steps=3, warmup=1, threads=threads)
        cpu = {"value": B * K * len(times) / sum(times), "unit": "frame-steps/s", "cores": threads, "kind": "port",
               "sample": f"{len(times)} diffusion steps (after 1 warm-up) of the same B={B},K={K} batch: oracle port of the "
                         f"reference path, fp32 torch CPU ops, {threads} threads"}

    # ------------------------------------------------------------------ secondary metric: train samples/s (BASELINE cfg2)
    train = None
    if not args.no_train:
        train = bench_train(dev, world, args.precision, "cfg2-train")
        train["cfg3"] = bench_train(dev, world, args.precision, "cfg3-train")

    if rank == 0:
        line = {"metric": "denoiser frame-steps/sec (sampling)", "value": value, "unit": "frame-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "ms_per_step_hot": hot_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "videos_per_gpu": B, "frames_per_stage": K, "n_obs": wl["n_obs"],
                           "latent": [C, S, S], **over, "sharding": f"video batch over {world} rank(s), no collective",
                           "l2": "flushed between timed steps (192 MB memset outside the event pairs)",
                           "weights": "random non-zero init (zero_module tensors re-randomised)"},
                "clocks": clk, "e2e": e2e, "gpu_launches": args.steps * (len(P.calls) + 1),
                "launches_per_step": len(P.calls) + 1, "roofline": roofline, "roofline_all_convs": all_convs,
                "step_roofline": step_roofline,
                "cpu_baseline": cpu, "train": train}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
