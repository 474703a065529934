"""cProfile of the host side of native training steps (the cfg2 step is launch/host-bound)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch as th
import bench
over, B, K, steps = bench.TRAIN_WORKLOADS["cfg2-train"]
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
for _ in range(5):
    step()
th.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
th.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
