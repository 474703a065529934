"""CPU: the oracle restatement reproduces the committed reference outputs (tests/golden/*.pt)."""
import pytest
import torch

from oracle import fdm_oracle as O


def _kw(inp):
    return dict(frame_indices=inp["frame_indices"], x0=inp["x0"], obs_mask=inp["obs_mask"], latent_mask=inp["latent_mask"])


@pytest.mark.parametrize("name", ["fwd_cfg1", "fwd_pad", "fwd_img64", "fwd_nc64"])
def test_forward_matches_reference(golden, name):
    g = golden(name)
    cfg = O.make_cfg(**g["over"])
    sd = O.init_state_dict(cfg, seed=1)
    inp = g["inputs"]
    with torch.no_grad():
        eps = O.unet_forward(sd, cfg, inp["x"], inp["x0"], g["model_t"], inp["frame_indices"], inp["obs_mask"], inp["latent_mask"])
    assert O.rel_l2(eps, g["eps"]) <= 2e-6
    assert float(g["eps"].abs().mean()) > 1e-2  # not the vacuous zero-init case


def test_param_contract():
    cfg = O.make_cfg(image_size=32, in_channels=4, num_channels=32, num_res_blocks=1)
    shapes = O.param_shapes(cfg)
    assert len(shapes) == 390
    assert sum(int(torch.Size(s).numel()) for s in shapes.values()) == 2_063_268
    cfg2 = O.make_cfg(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1)
    assert sum(int(torch.Size(s).numel()) for s in O.param_shapes(cfg2).values()) == 8_204_100


def test_sample_loop_matches_reference(golden):
    g = golden("sample_cfg1")
    cfg = O.make_cfg(**g["over"])
    sd = O.init_state_dict(cfg, seed=1)
    tab = O.Tables(cfg)
    assert tab.num_timesteps == g["num_timesteps"] == 4
    final = O.p_sample_loop(tab, sd, cfg, tuple(g["inputs"]["x0"].shape), _kw(g["inputs"]), g["noises"])
    assert O.rel_l2(final, g["final"]) <= 1e-5


def test_training_losses_match_reference(golden):
    g = golden("train_cfg1")
    cfg = O.make_cfg(**g["over"])
    sd = {k: v.requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=1).items()}
    inp = g["inputs"]
    terms = O.training_losses(O.Tables(cfg), sd, cfg, inp["x0"], g["t"], g["noise"], _kw(inp),
                              latent_mask=1 - inp["obs_mask"], eval_mask=inp["latent_mask"])
    terms["loss"].mean().backward()
    for k in ("loss", "mse", "eval-mse"):
        assert O.rel_l2(terms[k].detach(), g["terms"][k]) <= 2e-6
    for k, ref in g["grads"].items():
        assert O.rel_l2(sd[k].grad, ref) <= 1e-4, k
    # every parameter receives a gradient (DDP find_unused_parameters=False, train_util.py:124)
    assert all(v.grad is not None and float(v.grad.abs().sum()) > 0 for v in sd.values())


def test_attention_quirks():
    """residual on the normed input; two-group block-diagonal mask."""
    cfg = O.make_cfg(image_size=32, in_channels=4, num_channels=32, num_res_blocks=1)
    sd = O.init_state_dict(cfg, seed=2)
    p = "middle_block.1.temporal_attention"
    sd[p + ".proj_out.weight"].zero_()
    sd[p + ".proj_out.bias"].zero_()
    B, D, C, T = 1, 4, 64, 5
    x = torch.randn(B, D, C, T)
    temb = torch.randn(B * T, 128)
    fi = torch.tensor([[0, 3, 4, 9, 11]])
    mask = torch.tensor([[1., 1., 1., 0., 0.]])
    y, attn = O.rpe_attention(sd, p, x, temb, fi, mask, 4, True)
    gn = O.gn32(x.reshape(B * D, C, T), sd[p + ".norm.weight"], sd[p + ".norm.bias"]).view(B, D, C, T)
    assert torch.allclose(y, gn, atol=1e-6) and not torch.allclose(y, x, atol=1e-3)
    assert float(attn[..., :3, 3:].abs().max()) == 0 and float(attn[..., 3:, :3].abs().max()) == 0
    assert torch.allclose(attn.sum(-1), torch.ones_like(attn.sum(-1)), atol=1e-6)
