"""Whole-video sampling through the stage driver: B videos of 300 latent frames, hierarchy-2 (the reference's own
27-stage schedule, tests/golden/stages_hierarchy-2_T300.json generated from the live reference), n_obs = 36, K = 20.
Device-resident driver (video_sampler.sample_video_with_iterator) vs the reference's host-loop procedure
(scripts/video_sample.py:56-83 restated: CPU buffer, per-row gather, upload, sample, download, per-row scatter).
python tools/video_bench.py [respacing=50] [B=8]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch as th
import bench
from improved_diffusion import video_sampler

resp = sys.argv[1] if len(sys.argv) > 1 else "50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "stages_hierarchy-2_T300.json")))
over = dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000, timestep_respacing=resp)
dev = th.device("cuda:0")
model, diffusion, _ = bench.build_native(over, dev)
model.eval()
T, n_obs = fx["video_length"], fx["n_obs"]
batch = th.randn(B, T, 4, 32, 32, generator=th.Generator().manual_seed(0)).clamp(-1, 1).pin_memory()


class Replay:
    """the reference iterator's protocol over the committed stage list"""
    def __init__(self):
        self.i = 0

    def set_videos(self, v):
        pass

    def __iter__(self):
        return self

    def __next__(self):
        if self.i >= len(fx["stages"]):
            raise StopIteration
        o, l = fx["stages"][self.i]
        self.i += 1
        return [o] * B, [l] * B


@th.no_grad()
def host_loop():
    samples = th.zeros_like(batch)
    samples[:, :n_obs] = batch[:, :n_obs]
    for obs, lat in Replay():
        fi = th.cat([th.tensor(obs), th.tensor(lat)], dim=1).long()
        x0 = th.stack([samples[i, f] for i, f in enumerate(fi)], dim=0).clone()
        om = th.cat([th.ones_like(th.tensor(obs)), th.zeros_like(th.tensor(lat))], dim=1).view(B, -1, 1, 1, 1).float()
        lm = 1 - om
        x0, om, lm, fi = (t.to(dev) for t in (x0, om, lm, fi))
        out, _ = diffusion.p_sample_loop(model, x0.shape, clip_denoised=True, model_kwargs=dict(frame_indices=fi, x0=x0, obs_mask=om,
                                                                                              latent_mask=lm), latent_mask=lm)
        for i, li in enumerate(lat):
            samples[i, li] = out[i, -len(li):].cpu()
    return samples


def device_driver():
    out, used = video_sampler.sample_video_with_iterator(model, diffusion, batch, Replay(), n_obs, device=dev)
    assert len(used) == len(fx["stages"])
    return out


frame_steps = B * sum(len(o) + len(l) for o, l in fx["stages"]) * diffusion.num_timesteps
res = {}
for name, fn in (("device_resident_driver", device_driver), ("host_loop_reference_procedure", host_loop)):
    fn()  # warm-up: plans + graphs for both stage shapes (K = 20 and the ragged last stage)
    th.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    th.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert bool(th.isfinite(out).all())
    res[name] = dict(seconds=dt, frame_steps_per_s=frame_steps / dt)
print(json.dumps(dict(workload="cfg4-video", videos=B, video_length=T, stages=len(fx["stages"]), diffusion_steps=diffusion.num_timesteps,
                      frame_steps=frame_steps, **res)))
