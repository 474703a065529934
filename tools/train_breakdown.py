"""Where a native training step spends its time (CUDA events around the graph replays and the torch-side pieces)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch as th
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2-train"
over, B, K, steps = bench.TRAIN_WORKLOADS[wl]
dev = th.device("cuda:0")
model, diffusion, _ = bench.build_native(over, dev)
model.train()
from improved_diffusion.optim import FlatAdamW
opt = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=0.0)
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(over, B, K, 3, 4 * K, seed=1).items()}
t = th.randint(0, 1000, (B,), device=dev)
def step():
    terms = diffusion.training_losses(model, batch["x0"], t, model_kwargs=batch, latent_mask=1 - batch["obs_mask"], eval_mask=batch["latent_mask"])
    opt.zero_grad(set_to_none=True)
    terms["loss"].mean().backward()
    opt.step()
for _ in range(4):
    step()
th.cuda.synchronize()
P = next(iter(model.engine().train_plans.values()))
def timed(fn, n=20):
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    fn(); th.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); th.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(wl, "whole step            %.3f ms" % timed(step))
print("  forward graph replay  %.3f ms" % timed(lambda: P.graphs[("train", "fwd")].replay()))
print("  backward graph replay %.3f ms" % timed(lambda: P.graphs[("train", "bwd")].replay()))
print("  pgrad clone           %.3f ms" % timed(lambda: P.pgrad.clone()))
print("  optimizer step        %.3f ms" % timed(lambda: opt.step()))
def loss_only():
    eps = P.eps_view.clone().requires_grad_()
    ((eps - batch["x0"]) ** 2 * (1 - batch["obs_mask"])).mean(dim=(1, 2, 3, 4)).mean().backward()
print("  loss fwd+bwd (torch)  %.3f ms" % timed(loss_only))
import time
th.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter(); th.cuda.synchronize(); t2 = time.perf_counter()
print("  host time per step (launch side) %.3f ms, drained %.3f ms" % ((t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))
