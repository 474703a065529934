"""Helper run as a SUBPROCESS by tests/test_gpu_reference.py on the B200 box.

The reference's own TrainLoop (oracle/_ref/improved_diffusion/train_util.py, UNMODIFIED: sample_all_masks, forward_backward through
DistributedDataParallel on CUDA, optimize_normal with torch AdamW, update_ema, log_step) runs

  (A) over the reference's model + diffusion (PyTorch eager fp32, TF32 off)   — package alias `fdm_ref_improved_diffusion`
  (B) over THIS repo's model + diffusion on the native forward + backward kernel schedules (FDM_TRAIN_ENGINE=native, fp32 mode)
      — the mixed package: hot-path modules from this repo, train_util & co. from the reference (FDM_REFERENCE_PATH)

from the same weights, data stream and seeds.  Compared: the gradients after the first forward_backward (rel-L2 per parameter),
the logged losses of two full run_step()s, and that both loops end with finite parameters.  Post-Adam parameters are NOT compared
element-wise: the first Adam step is lr*sign(g), so gradient elements at rounding-noise level legitimately flip.
mpi4py / blobfile are the single-process stand-ins of oracle/stubs; wandb is stubbed (no network)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch as th  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import fdm_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

logged = []
sys.modules["wandb"] = types.SimpleNamespace(log=lambda d, **k: logged.append(dict(d)), Video=lambda *a, **k: None,
                                             run=types.SimpleNamespace(id="test"), init=lambda **k: None)
R.add_stubs()
R.enable_mixed_package()

import improved_diffusion  # noqa: E402
from improved_diffusion import train_util as mixed_train_util, unet as native_unet  # noqa: E402
from improved_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults  # noqa: E402

assert "_b200" in native_unet.__file__ and os.path.join("oracle", "_ref") in mixed_train_util.__file__, \
    (native_unet.__file__, mixed_train_util.__file__)
ref_train_util = R.ref_module("train_util")

assert th.cuda.is_available()
th.backends.cudnn.allow_tf32 = False
th.backends.cuda.matmul.allow_tf32 = False
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29631")
os.environ["DIFFUSION_TRAINING_TEST"] = "1"
th.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=th.device("cuda", 0))

over = dict(image_size=32, in_channels=4, num_channels=32, num_res_blocks=1, diffusion_steps=32)
PIXEL = dict(diffusion_space="pixel", pre_encoded=False, pre_encoded_stats_dict=None)
cfg = O.make_cfg(**over)
sd = O.init_state_dict(cfg, seed=1)


def data():
    g = th.Generator().manual_seed(1)
    while True:
        yield th.randn(2, 12, 4, 32, 32, generator=g).clamp(-1, 1), {}


def make_loop(tu, model, diffusion):
    tu.TrainLoop.save = lambda self: None
    return tu.TrainLoop(model=model, diffusion=diffusion, data=data(), batch_size=2, microbatch=-1, lr=1e-4, ema_rate="0.9999",
                        log_interval=1, save_interval=10 ** 9, resume_checkpoint="", use_fp16=False,
                        diffusion_space_kwargs=dict(PIXEL), fp16_scale_growth=1e-3, schedule_sampler=None, weight_decay=0.0,
                        lr_anneal_steps=0, sample_interval=None, pad_with_random_frames=True, max_frames=5,
                        enc_dec_chunk_size=10, args=types.SimpleNamespace(resume_id="", T=12))


def run(tu, model, diffusion, tag):
    model.load_state_dict(sd, strict=True)
    model.cuda().train()
    loop = make_loop(tu, model, diffusion)
    assert loop.use_ddp, "on CUDA the reference wraps the model in DistributedDataParallel (train_util.py:118-125)"
    th.manual_seed(5)
    np.random.seed(5)
    loop.forward_backward()
    th.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu().clone() for k, p in model.named_parameters()}
    assert all(p.grad is not None for p in model.parameters()), "every parameter must receive a gradient (find_unused_parameters=False)"
    n0 = len(logged)
    loop.optimize_normal()
    loop.log_step()
    tu.logger.dumpkvs()
    loop.step += 1
    loop.run_step()  # a complete second step through the reference's own run_step (train_util.py:267-275)
    tu.logger.dumpkvs()
    th.cuda.synchronize()
    assert all(bool(th.isfinite(p).all()) for p in model.parameters())
    print(f"{tag}: two steps done, |grads| = {float(th.stack([g.norm() for g in grads.values()]).norm()):.4e}")
    return grads, logged[n0:]


d = model_and_diffusion_defaults()
d.update(over)
d["diffusion_space_kwargs"] = dict(PIXEL)
ref_model, ref_diffusion = R.create_reference(over)
g_ref, log_ref = run(ref_train_util, ref_model, ref_diffusion, "reference TrainLoop over the reference model")

assert os.environ.get("FDM_TRAIN_ENGINE", "native") == "native" and os.environ.get("FDM_ALLOW_TORCH_TRAIN") != "1"
model, diffusion = create_model_and_diffusion(**d)
model.precision = "fp32"
g_nat, log_nat = run(mixed_train_util, model, diffusion, "reference TrainLoop over the NATIVE model")
plans = model.engine("fp32").train_plans
assert len(plans) >= 1, "the native training plan was never built: the step did not run on the kernel schedules"

floor = 5e-2 * float(th.stack([v.double().norm() for v in g_ref.values()]).median())
errs = sorted(((float((g_nat[k].double() - g_ref[k].double()).norm() / max(float(g_ref[k].double().norm()), floor)), k)
               for k in g_ref), reverse=True)
print("worst gradient rel-L2 (native vs reference, fp32):", errs[:3])
assert errs[0][0] <= 2e-3, errs[:5]


def losses(entries):
    return [float(e["loss"]) for e in entries if "loss" in e]


la, lb = losses(log_ref), losses(log_nat)
print("logged losses: reference", la, "native", lb)
assert len(la) == len(lb) >= 1
# step 1 losses see identical weights; step 2 follows one Adam step whose sign(g) updates may differ on noise-level elements
assert abs(la[0] - lb[0]) <= 1e-4 * abs(la[0]), (la, lb)
assert abs(la[-1] - lb[-1]) <= 2e-2 * abs(la[-1]), (la, lb)
print("TRAINLOOP_GPU_OK")
dist.destroy_process_group()
