"""Run under torchrun with >= 2 GPUs (gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 ... tests/multigpu_ddp_overlap.py).

sharding.FlatGradDataParallel over NCCL: the BUCKETED gradient exchange overlapped with the native backward (default) against the
single allreduce after the whole backward (overlap=False) — same averaged gradients on every rank, and the step time of both
(CUDA events, max over ranks).  Prints one JSON line per workload on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch as th  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from improved_diffusion import sharding  # noqa: E402
from improved_diffusion.optim import FlatAdamW  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local)
dev = th.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
os.environ["FDM_TRAIN_ENGINE"] = "native"


def run(workload, overlap, steps):
    over, B, K, _ = bench.TRAIN_WORKLOADS[workload]
    model, diffusion, _ = bench.build_native(over, dev)
    model.precision = "bf16"
    model.train()
    # overlap=None: no data-parallel wrapper at all — the local gradient, averaged by hand below (the reference for both modes)
    net = sharding.FlatGradDataParallel(model, overlap=overlap) if overlap is not None else model
    opt = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=0.0, model=model)
    batch = {k: v.to(dev) for k, v in bench.synthetic_batch(over, B, K, 3, 4 * K, seed=1 + rank).items()}
    g = th.Generator(device=dev).manual_seed(rank)

    def step(update=True):
        t = th.randint(0, diffusion.num_timesteps, (B,), device=dev, generator=g)
        terms = diffusion.training_losses(net, batch["x0"], t, model_kwargs=batch, latent_mask=1 - batch["obs_mask"],
                                          eval_mask=batch["latent_mask"])
        opt.zero_grad()
        terms["loss"].mean().backward()
        if update:
            opt.step()
    g.manual_seed(100 + rank)
    for _ in range(3):
        step(update=False)  # warm-up / graph capture without moving the weights: both modes see identical parameters below
    g.manual_seed(200 + rank)
    th.manual_seed(1000 + rank)  # q_sample's noise comes from the global CUDA generator: identical in the three runs
    step(update=False)
    th.cuda.synchronize()
    grad = opt.flat_g.clone()
    if overlap is None:
        dist.all_reduce(grad)
        grad /= world
        P = next(iter(model.engine().train_plans.values()))
        return grad, 0.0, len(P.grad_buckets)
    for _ in range(2):
        step()
    th.cuda.synchronize()
    dist.barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    th.cuda.synchronize()
    ms = th.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=th.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    P = next(iter(model.engine().train_plans.values()))
    nb = len(P.grad_buckets)
    del opt, net, model
    th.cuda.empty_cache()
    return grad, float(ms.item()), nb


for workload, steps in (("cfg2-train", 30), ("cfg3-train", 8)):
    gref, _, _ = run(workload, None, steps)
    g1, ms1, nb = run(workload, True, steps)
    g0, ms0, _ = run(workload, False, steps)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    if rank == 0:
        print(f"{workload}: vs hand-averaged local gradients: overlapped {rel(g1, gref):.3e}, single allreduce {rel(g0, gref):.3e}", flush=True)
    err = max(rel(g1, gref), rel(g0, gref))
    # every rank must hold the same averaged gradient
    ref = g1.clone()
    dist.broadcast(ref, 0)
    same = float((g1.double() - ref.double()).norm() / ref.double().norm())
    worst = th.tensor([err, same], device=dev, dtype=th.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": workload, "n_gpus": world, "buckets": nb, "ms_per_step_overlapped": ms1, "ms_per_step_single_allreduce": ms0,
                          "grad_rel_l2_overlap_vs_single": float(worst[0]), "grad_rel_l2_across_ranks": float(worst[1])}), flush=True)
    assert float(worst[0]) <= 2e-2 and float(worst[1]) == 0.0, worst  # bf16 run-to-run noise (atomics) vs exact agreement across ranks
if rank == 0:
    print("DDP_OVERLAP_OK")
dist.destroy_process_group()
