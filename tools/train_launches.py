"""One native training step (forward + backward schedules, eager launches) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`:  python tools/train_launches.py <cfg> [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch as th
import bench
CFG = {"cfg2": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000), 1, 5),
       "cfg4": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000), 8, 20),
       "cfg3": (dict(image_size=128, in_channels=3, num_channels=128, num_res_blocks=1, diffusion_steps=1000), 2, 20)}
over, B, K = CFG[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = th.device("cuda:0")
model, diffusion, _ = bench.build_native(over, dev)
model.precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
model.train()
os.environ["FDM_NO_GRAPH"] = "1"
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(over, B, K, 3, 4 * K, seed=1).items()}
t = th.randint(0, 1000, (B,), device=dev)

def step():
    terms = diffusion.training_losses(model, batch["x0"], t, model_kwargs=batch, latent_mask=1 - batch["obs_mask"], eval_mask=batch["latent_mask"])
    model.zero_grad(set_to_none=True)
    terms["loss"].mean().backward()
for _ in range(2):
    step()
th.cuda.synchronize()
th.cuda.profiler.start()
step()
th.cuda.synchronize()
th.cuda.profiler.stop()
