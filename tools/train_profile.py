"""Per-entry-point CUDA-event time of one training step's forward and backward schedules (eager launches, one event pair per
call): python tools/train_profile.py <cfg> [precision].  cfg: cfg2 | cfg3 | cfg4."""
import collections, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch as th
import bench
from improved_diffusion import _native as N_

CFG = {"cfg2": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000), 1, 5),
       "cfg2b8": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000), 8, 5),
       "cfg4": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000), 8, 20),
       "cfg5": (dict(image_size=64, in_channels=4, num_channels=128, num_res_blocks=1, diffusion_steps=1000), 2, 40),
       "cfg3": (dict(image_size=128, in_channels=3, num_channels=128, num_res_blocks=1, diffusion_steps=1000), 2, 20)}
over, B, K = CFG[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = th.device("cuda:0")
model, diffusion, _ = bench.build_native(over, dev)
model.precision = prec
model.train()
os.environ["FDM_NO_GRAPH"] = "1"
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(over, B, K, 3, 4 * K, seed=1).items()}
t = th.randint(0, 1000, (B,), device=dev)
for _ in range(2):
    terms = diffusion.training_losses(model, batch["x0"], t, model_kwargs=batch, latent_mask=1 - batch["obs_mask"], eval_mask=batch["latent_mask"])
    model.zero_grad(set_to_none=True)
    terms["loss"].mean().backward()
th.cuda.synchronize()
P = next(iter(model.engine().train_plans.values()))
s = C.c_void_p(th.cuda.current_stream().cuda_stream)
print(f"{sys.argv[1:]}: arena {P.arena_bytes / 1e9:.2f} GB, fwd flops {P.flops / 1e9:.1f} G, bwd flops {P.bflops / 1e9:.1f} G")
for which, calls, ops in (("forward", P.calls, P.ops), ("backward", P.bcalls, P.bops)):
    agg = collections.defaultdict(lambda: [0.0, 0])
    total = 0.0
    for rep in range(3):
        for (name, fn, ref), (_, _, f) in zip(calls, ops):
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(ref, s)
            e1.record()
            e1.synchronize()
            assert rc == 0, name
            if rep == 2:
                key = name
                if name == "fdm_conv":
                    key += f"[k{f['ksize']} {'TC' if f['engine'] else 'SIMT'}]"
                if name == "fdm_conv_wgrad":
                    key += f"[k{f['ksize']}]"
                ms = e0.elapsed_time(e1)
                agg[key][0] += ms
                agg[key][1] += 1
                total += ms
    print(f"--- {which}: {len(calls)} launches, sum {total:.3f} ms")
    for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {ms:9.3f} ms {100 * ms / total:5.1f}%  n={n:3d}  {k}")
