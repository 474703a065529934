"""
Generate the golden fixtures under tests/golden/ from the LIVE, UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case it (1) builds the reference model via the reference's own
create_model_and_diffusion, (2) loads the deterministic non-zero weights of
oracle.fdm_oracle.init_state_dict through the reference's load_state_dict(strict=True) — which also
pins the state-dict key/shape contract — (3) runs the reference on seeded synthetic inputs and
(4) stores inputs + reference outputs.  It then asserts that the oracle restatement reproduces the
reference on the same inputs (rel-L2 <= 2e-6) before writing, so a stale oracle cannot be pinned.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FDM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import fdm_oracle as O  # noqa: E402
from improved_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults  # noqa: E402
import improved_diffusion  # noqa: E402

assert improved_diffusion.__file__.startswith(REF), improved_diffusion.__file__

PIXEL = dict(diffusion_space="pixel", pre_encoded=False, pre_encoded_stats_dict=None)


def build_ref(over):
    d = model_and_diffusion_defaults()
    d.update(over)
    d["diffusion_space_kwargs"] = dict(PIXEL)
    model, diffusion = create_model_and_diffusion(**d)
    cfg = O.make_cfg(**over)
    sd = O.init_state_dict(cfg, seed=1)
    assert list(model.state_dict().keys()) == list(sd.keys()), "state-dict key order/contents differ"
    model.load_state_dict(sd, strict=True)
    return model, diffusion, cfg, sd


def kw_of(inp):
    return dict(frame_indices=inp["frame_indices"], x0=inp["x0"], obs_mask=inp["obs_mask"], latent_mask=inp["latent_mask"])


def check(name, a, b, tol=2e-6):
    e = O.rel_l2(a, b)
    print(f"  oracle vs reference [{name}]: rel-L2 = {e:.3e}")
    assert e <= tol, (name, e)


def fwd_case(name, over, B, T, n_obs, pad_rows=(), t_val=None, seed=0):
    print(name)
    model, diffusion, cfg, sd = build_ref(over)
    model.eval()
    inp = O.synthetic_inputs(cfg, B, T, n_obs, seed=seed, pad_rows=pad_rows)
    n = diffusion.num_timesteps
    t = torch.tensor([(7 * (i + 1)) % n for i in range(B)]) if t_val is None else torch.tensor([t_val] * B)
    tab = O.Tables(cfg)
    ts = O.model_timesteps(tab, t)
    with torch.no_grad():
        eps_ref, _ = model(inp["x"], timesteps=ts, **kw_of(inp))
        taps = {}
        eps_or = O.unet_forward(sd, cfg, inp["x"], inp["x0"], ts, inp["frame_indices"], inp["obs_mask"], inp["latent_mask"], taps=taps)
    check("eps", eps_or, eps_ref)
    assert float(eps_ref.abs().mean()) > 1e-2, "vacuous parity: eps ~ 0"
    torch.save(dict(over=over, B=B, T=T, n_obs=n_obs, pad_rows=list(pad_rows), seed=seed, t=t, model_t=ts,
                    eps=eps_ref.clone(), inputs=inp), os.path.join(HERE, name + ".pt"))


def sample_case(name, over, B, T, n_obs, seed=0):
    print(name)
    model, diffusion, cfg, sd = build_ref(over)
    model.eval()
    inp = O.synthetic_inputs(cfg, B, T, n_obs, seed=seed)
    shape = tuple(inp["x0"].shape)
    n = diffusion.num_timesteps
    torch.manual_seed(1234)
    noises = [torch.randn(*shape)] + [torch.randn(*shape) for _ in range(n)]
    torch.manual_seed(1234)  # the reference draws th.randn(*shape) then one randn_like per step
    final, _ = diffusion.p_sample_loop(model, shape, clip_denoised=True, model_kwargs=kw_of(inp), latent_mask=inp["latent_mask"])
    tab = O.Tables(cfg)
    trace = []
    final_or = O.p_sample_loop(tab, sd, cfg, shape, kw_of(inp), noises, trace=trace)
    check("final sample", final_or, final, tol=1e-5)
    # per-step references for the first and last step (single-step parity on the GPU side)
    torch.save(dict(over=over, B=B, T=T, n_obs=n_obs, seed=seed, noises=noises, final=final.clone(),
                    step_samples=[o["sample"].clone() for o in trace], step_eps=[o["eps"].clone() for o in trace],
                    inputs=inp, num_timesteps=n), os.path.join(HERE, name + ".pt"))


def train_case(name, over, B, T, n_obs, pad_rows=(), seed=0):
    print(name)
    model, diffusion, cfg, sd = build_ref(over)
    model.train()
    inp = O.synthetic_inputs(cfg, B, T, n_obs, seed=seed, pad_rows=pad_rows)
    g = torch.Generator().manual_seed(99)
    noise = torch.randn(inp["x0"].shape, generator=g)
    n = diffusion.num_timesteps
    t = torch.tensor([(11 * (i + 1)) % n for i in range(B)])
    kw = dict(frame_indices=inp["frame_indices"], obs_mask=inp["obs_mask"], latent_mask=inp["latent_mask"], x0=inp["x0"])
    # train_util.py:298-310: latent_mask=(1-obs_mask), eval_mask=latent_mask
    terms = diffusion.training_losses(model, inp["x0"], t, model_kwargs=kw, noise=noise,
                                      latent_mask=1 - inp["obs_mask"], eval_mask=inp["latent_mask"])
    terms["loss"].mean().backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    assert all(g is not None for g in grads.values())
    # oracle
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tab = O.Tables(cfg)
    terms_or = O.training_losses(tab, sdg, cfg, inp["x0"], t, noise, kw, latent_mask=1 - inp["obs_mask"], eval_mask=inp["latent_mask"])
    terms_or["loss"].mean().backward()
    for k in ("loss", "mse", "eval-mse"):
        check(k, terms_or[k], terms[k].detach())
    worst = max(O.rel_l2(sdg[k].grad, grads[k]) for k in grads)
    print(f"  oracle vs reference [worst param-grad rel-L2 over {len(grads)} tensors]: {worst:.3e}")
    assert worst <= 1e-4
    keep = ["input_blocks.0.0.weight", "time_embed.0.weight", "out.2.weight", "out.2.bias",
            "middle_block.1.temporal_attention.rpe_k.rpe_net.out.weight",
            "middle_block.1.temporal_attention.qkv.weight", "middle_block.1.spatial_attention.proj_out.weight",
            "output_blocks.2.0.skip_connection.weight", "input_blocks.1.0.emb_layers.1.weight",
            "input_blocks.1.0.out_layers.0.weight"]
    torch.save(dict(over=over, B=B, T=T, n_obs=n_obs, pad_rows=list(pad_rows), seed=seed, t=t, noise=noise, inputs=inp,
                    terms={k: v.detach().clone() for k, v in terms.items()},
                    grad_norms={k: float(v.norm()) for k, v in grads.items()},
                    grads={k: grads[k] for k in keep}), os.path.join(HERE, name + ".pt"))


def attn_case(name, over, B, T, n_obs, seed=0):
    """p_sample_loop(return_attn_weights=True): the per-quartile attention-map averages that TrainLoop.log_samples logs."""
    print(name)
    model, diffusion, cfg, sd = build_ref(over)
    model.eval()
    inp = O.synthetic_inputs(cfg, B, T, n_obs, seed=seed)
    shape = tuple(inp["x0"].shape)
    n = diffusion.num_timesteps
    torch.manual_seed(4321)
    noises = [torch.randn(*shape) for _ in range(n + 1)]
    torch.manual_seed(4321)
    final, attns = diffusion.p_sample_loop(model, shape, clip_denoised=True, model_kwargs=kw_of(inp), latent_mask=inp["latent_mask"],
                                           return_attn_weights=True)
    assert attns, "reference returned no attention maps"
    torch.save(dict(over=over, B=B, T=T, n_obs=n_obs, seed=seed, noises=noises, final=final.clone(), inputs=inp,
                    attns={k: v.clone() for k, v in attns.items()}), os.path.join(HERE, name + ".pt"))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    small = dict(image_size=32, in_channels=4, num_channels=32, num_res_blocks=1, diffusion_steps=32)
    fwd_case("fwd_cfg1", small, B=1, T=5, n_obs=3)
    fwd_case("fwd_pad", small, B=2, T=4, n_obs=2, pad_rows=(1,), seed=3)
    fwd_case("fwd_img64", dict(image_size=64, in_channels=3, num_channels=32, num_res_blocks=1, diffusion_steps=1000),
             B=1, T=2, n_obs=1, seed=5)
    fwd_case("fwd_nc64", dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000),
             B=1, T=5, n_obs=3, seed=7)
    sample_case("sample_cfg1", dict(small, timestep_respacing="4"), B=1, T=5, n_obs=3)
    train_case("train_cfg1", small, B=2, T=4, n_obs=2, pad_rows=(1,), seed=11)
    attn_case("attn_cfg1", dict(small, timestep_respacing="4"), B=2, T=3, n_obs=1, seed=13)
    print("golden fixtures written to", HERE)
