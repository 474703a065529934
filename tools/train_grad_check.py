"""Per-parameter comparison of the native backward schedule with torch.autograd over the PyTorch expression (GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch
from oracle import fdm_oracle as O
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import build, cuda_kw

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
over = dict(image_size=32, in_channels=4, num_channels=32, num_res_blocks=1, diffusion_steps=32)
model, diffusion, cfg, sd = build(over, prec)
model.train()
inp = O.synthetic_inputs(cfg, 2, 5, 2, seed=5, video_len=60, pad_rows=(1,))
g = torch.Generator().manual_seed(11)
t = torch.randint(0, diffusion.num_timesteps, (2,), generator=g)
noise = torch.randn(inp["x0"].shape, generator=g)

def run(engine):
    os.environ["FDM_TRAIN_ENGINE"] = engine
    model.zero_grad(set_to_none=True)
    terms = diffusion.training_losses(model, inp["x0"].cuda(), t.cuda(), model_kwargs=cuda_kw(inp), noise=noise.cuda(),
                                      latent_mask=inp["latent_mask"].cuda(), eval_mask=inp["latent_mask"].cuda())
    terms["loss"].mean().backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().double().cpu() for k, p in model.named_parameters()}

a, n = run("autograd"), run("native")
scale = torch.stack([v.norm() for v in a.values()]).median()
rows = sorted(((float((n[k] - a[k]).norm() / a[k].norm().clamp_min(1e-30)), float(a[k].norm()), float(n[k].norm()), k) for k in a), reverse=True)
print("median grad norm", float(scale))
for e, na, nn_, k in rows[:25]:
    print(f"{e:10.3e}  |auto|={na:10.3e} |native|={nn_:10.3e}  {k}")
