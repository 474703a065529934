"""Whole training step through train_step.NativeTrainStep with HOST video batches (mask sampling, gather, pinned upload,
forward, backward, gradient norm, AdamW + EMA, one log read per step), native stack vs the torch stack (torch.autograd +
torch.optim.AdamW + per-tensor EMA) of the same class.  Usage: python tools/train_step_bench.py [cfg2|cfg3] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "latent-flexible-video-diffusion-modeling_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from improved_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults  # noqa: E402
from improved_diffusion.train_step import NativeTrainStep  # noqa: E402

CFG = {"cfg2": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1), 1, 5, 60),
       "cfg3": (dict(image_size=128, in_channels=3, num_channels=128, num_res_blocks=1), 2, 20, 60)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    over, B, K, T = CFG[name]
    out = {"workload": f"{name}: B={B} videos of {T} frames on the host, max_frames={K}", "steps": steps}
    arms = (("native", "flat", "native"), ("torch", "torch", "autograd"))
    if os.environ.get("FDM_TSB_ARMS") == "native":
        arms = arms[:1]
    for arm, optimizer, engine in arms:
        os.environ["FDM_TRAIN_ENGINE"] = engine
        os.environ["FDM_ALLOW_TORCH_TRAIN"] = "1"
        d = model_and_diffusion_defaults()
        d.update(over, diffusion_steps=1000,
                 diffusion_space_kwargs=dict(diffusion_space="pixel", pre_encoded=False, pre_encoded_stats_dict=None))
        torch.manual_seed(0)
        model, diffusion = create_model_and_diffusion(**d)
        model.to("cuda").train()
        runner = NativeTrainStep(model, diffusion, lr=1e-4, max_frames=K, ema_rate="0.9999", optimizer=optimizer)
        torch.manual_seed(1)
        np.random.seed(1)
        g = torch.Generator().manual_seed(2)
        C, S = over["in_channels"], over["image_size"]
        pool = [torch.randn(B, T, C, S, S, generator=g).clamp(-1, 1) for _ in range(4)]
        if os.environ.get("FDM_TSB_DEVICE_POOL") == "1":  # videos already on the device: gather there, no upload
            pool = [v.cuda() for v in pool]
        n_arm = steps if arm == "native" else max(5, steps // 3)
        for i in range(5):
            runner.run_step(pool[i % 4], pool[(i + 1) % 4])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n_arm):
            log = runner.run_step(pool[i % 4], pool[(i + 1) % 4])
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n_arm * 1e3
        ms_defer = None
        if arm == "native":  # log reads left in flight: the host side of step i+1 overlaps the GPU work of step i
            t0 = time.perf_counter()
            for i in range(n_arm):
                runner.run_step(pool[i % 4], pool[(i + 1) % 4], defer=True)
            runner.flush()
            torch.cuda.synchronize()
            ms_defer = round((time.perf_counter() - t0) / n_arm * 1e3, 3)
        th0 = time.perf_counter()
        for i in range(20):
            runner_masks = __import__("improved_diffusion.train_step", fromlist=["x"]).sample_all_masks(
                pool[0], pool[1], max_frames=K)
        host_ms = (time.perf_counter() - th0) / 20 * 1e3
        out[arm] = {"ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 2), "ms_per_step_deferred_logs": ms_defer, "mask_and_gather_ms": round(host_ms, 3),
                    "loss": log["loss"], "grad_norm": log["grad_norm"]}
        del runner, model
        torch.cuda.empty_cache()
    if "torch" in out:
        out["speedup"] = round(out["torch"]["ms_per_step"] / out["native"]["ms_per_step"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
