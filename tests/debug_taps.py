"""Debug helper (not a test): per-layer rel-L2 of the GPU plan vs the oracle.  FDM_DEBUG_TAPS=1 python tests/debug_taps.py B T n_obs [precision]"""
import os, sys
os.environ["FDM_DEBUG_TAPS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200")]
import torch
from oracle import fdm_oracle as O
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import build, cuda_kw

B, T, n_obs = map(int, sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
over = dict(image_size=32, in_channels=4, num_channels=32, num_res_blocks=1, diffusion_steps=1000)
model, diffusion, cfg, sd = build(over, prec)
inp = O.synthetic_inputs(cfg, B, T, n_obs, seed=B * 10 + T, video_len=300)
ts = torch.tensor([37.0 * (b + 1) for b in range(B)])
taps = {}
with torch.no_grad():
    ref = O.unet_forward(sd, cfg, inp["x"], inp["x0"], ts, inp["frame_indices"], inp["obs_mask"], inp["latent_mask"], taps=taps)
    eps, _ = model(inp["x"].cuda(), timesteps=ts.cuda(), **cuda_kw(inp))
P = next(iter(model.engine().plans.values()))
for name, t in taps.items():
    key = name
    if key not in P.taps:
        key = name.rsplit(".", 1)[0] if name.rsplit(".", 1)[0] in P.taps else None
    if key is None:
        print("no tap for", name); continue
    print(f"{name:40s} {O.rel_l2(P.tap(key).cpu(), t):.3e}")
print("eps", O.rel_l2(eps.cpu(), ref))
