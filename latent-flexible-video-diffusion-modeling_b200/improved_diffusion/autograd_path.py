"""
Differentiable forward of UNetVideoModel for TRAINING (reference unet.py:428-464, rpe.py:133-174).

On a CUDA device training runs on the native forward + backward kernel schedules (engine._DenoiserFn); this module is the
PyTorch-autograd expression of the SAME network over the SAME parameters, kept for three callers: (1) CPU tensors (the host-side
tests, the drop-in check under the reference's TrainLoop, the 2-rank gloo DDP test), (2) the A/B arm of the training benchmark and
of the gradient-parity tests (`FDM_TRAIN_ENGINE=autograd`), (3) `return_attn_weights=True` (attention-map logging needs the
materialised attention matrices).  Every parameter receives a gradient (DDP find_unused_parameters=False, train_util.py:124).

`model.precision == "bf16"` runs the convolutions / linears under torch.autocast(bfloat16) (GroupNorm and softmax
stay fp32 as in the reference's GroupNorm32 / softmax(w.float())); "fp32" disables TF32 so the 1e-4 contract holds.
"""
import contextlib

import torch as th
import torch.nn.functional as F

from .nn import timestep_embedding


def _gn(norm, x):
    return F.group_norm(x.float(), norm.num_groups, norm.weight, norm.bias, norm.eps).type(x.dtype)


def _silu(x):
    return x * th.sigmoid(x)


def _res_block(rb, x, emb):
    h = rb.in_layers[2](_silu(_gn(rb.in_layers[0], x)))
    e = rb.emb_layers[1](_silu(emb)).type(h.dtype)[:, :, None, None]
    if rb.use_scale_shift_norm:
        scale, shift = th.chunk(e, 2, dim=1)  # scale = first half (unet.py:199-203)
        h = _silu(_gn(rb.out_layers[0], h) * (1 + scale) + shift)
    else:
        h = _silu(_gn(rb.out_layers[0], h + e))
    h = rb.out_layers[3](rb.out_layers[2](h))
    return rb.skip_connection(x) + h


def _rpe_table(net, temb, dist, heads):
    """RPENet (rpe.py:20-31): R[b,t,s,h,f] from the time embedding of query frame t and the index delta fi[t]-fi[s]."""
    feats = th.stack([th.log(1 + dist.clamp(min=0)), th.log(1 + (-dist).clamp(min=0)), (dist == 0).float()], dim=-1)
    B, T, _ = dist.shape
    C = net.out.weight.shape[0]
    e = net.embed_diffusion_time(temb).view(B, T, 1, C) + net.embed_distances(feats)
    return net.out(F.silu(e)).view(B, T, T, heads, C // heads)


def _attention(att, x, temb, frame_indices, attn_mask, collect=None):
    """RPEAttention._forward over the last axis of x [B, D, C, T] (rpe.py:133-174), quirks included: the residual is
    added to the GroupNorm-ed input; GroupNorm statistics span (C/32 x T); the mask is two-group block-diagonal."""
    B, D, C, T = x.shape
    H = att.num_heads
    xn = _gn(att.norm, x.reshape(B * D, C, T)).view(B, D, C, T).permute(0, 1, 3, 2)  # B D T C
    qkv = att.qkv(xn).reshape(B, D, T, 3, H, C // H).permute(3, 0, 1, 4, 2, 5)        # 3 B D H T F
    q, k, v = qkv[0] * att.scale, qkv[1], qkv[2]
    w = q @ k.transpose(-2, -1)
    if att.has_rpe:
        dist = frame_indices.unsqueeze(-1) - frame_indices.unsqueeze(-2)
        w = w + th.einsum("bdhtf,btshf->bdhts", q, _rpe_table(att.rpe_k.rpe_net, temb, dist, H).type(q.dtype))
        w = w + th.einsum("bdhtf,btshf->bdhts", k * att.scale,
                          _rpe_table(att.rpe_q.rpe_net, temb, dist, H).type(q.dtype)).transpose(-1, -2)
    if attn_mask is not None:
        m = attn_mask.view(B, 1, T)
        allowed = m * m.transpose(1, 2) + (1 - m) * (1 - m.transpose(1, 2))
        w = w.masked_fill((allowed == 0).view(B, 1, 1, T, T), float("-inf"))
    p = th.softmax(w.float(), dim=-1).type(w.dtype)
    out = p @ v
    if att.has_rpe:
        out = out + th.einsum("bdhts,btshf->bdhtf", p, _rpe_table(att.rpe_v.rpe_net, temb, dist, H).type(p.dtype))
    if collect is not None:  # attention-map logging (rpe.py:128-130): mean over heads, absolute value
        collect.append(p.detach().reshape(B * D, -1, T, T).mean(dim=1).abs())
    out = att.proj_out(out.permute(0, 1, 3, 2, 4).reshape(B, D, T, C))
    return (xn + out).permute(0, 1, 3, 2)


def _factorized_attention(fab, x, temb, attn_mask, T, frame_indices, attns=None):
    BT, C, H, W = x.shape
    B = BT // T
    x = x.view(B, T, C, H, W).permute(0, 3, 4, 2, 1).reshape(B, H * W, C, T)
    x = _attention(fab.temporal_attention, x, temb, frame_indices, attn_mask, None if attns is None else attns["temporal"])
    x = x.reshape(B, H, W, C, T).permute(0, 4, 3, 1, 2).reshape(B, T, C, H * W)
    x = _attention(fab.spatial_attention, x, temb, None, None, None if attns is None else attns["spatial"])
    return x.reshape(BT, C, H, W)


def _run(stage, h, emb, attn_mask, T, frame_indices, attns=None):
    from .unet import Downsample, FactorizedAttentionBlock, ResBlock, Upsample
    for layer in stage:
        if isinstance(layer, ResBlock):
            h = _res_block(layer, h, emb)
        elif isinstance(layer, FactorizedAttentionBlock):
            h = _factorized_attention(layer, h, emb, attn_mask, T, frame_indices, attns)
        elif isinstance(layer, Downsample):
            h = layer.op(h)
        elif isinstance(layer, Upsample):
            h = layer.conv(F.interpolate(h, scale_factor=2, mode="nearest"))
        else:  # stem conv
            h = layer(h)
    return h


def differentiable_forward(model, x, x0, timesteps, frame_indices, obs_mask, latent_mask, attns=None):
    """`attns`: optional {"spatial": [], "temporal": [], "mixed": []} filled with the per-block attention maps exactly as
    UNetVideoModel.forward(return_attn_weights=True) does upstream (unet.py:451-463) — the logging path of
    TrainLoop.log_samples (train_util.py:451-463); it needs materialised attention matrices, so it runs here, not on the
    fused kernels."""
    B, T, C, H, W = x.shape
    bf16 = model.precision == "bf16" and x.is_cuda
    ctx = th.autocast("cuda", dtype=th.bfloat16) if bf16 else contextlib.nullcontext()
    tf32 = (th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32)
    if not bf16:
        th.backends.cudnn.allow_tf32 = False
        th.backends.cuda.matmul.allow_tf32 = False
    try:
        with ctx:
            t = timesteps.view(B, 1).expand(B, T).reshape(B * T)
            attn_mask = (obs_mask + latent_mask).clip(max=1).flatten(start_dim=2).squeeze(dim=2)
            ind = th.ones_like(x[:, :, :1]) * obs_mask
            h = th.cat([x * (1 - obs_mask) + x0 * obs_mask, ind], dim=2).reshape(B * T, C + 1, H, W)
            emb = model.time_embed(timestep_embedding(t, model.model_channels))
            hs = []
            for stage in model.input_blocks:
                h = _run(stage, h, emb, attn_mask, T, frame_indices, attns)
                hs.append(h)
            h = _run(model.middle_block, h, emb, attn_mask, T, frame_indices, attns)
            for stage in model.output_blocks:
                h = _run(stage, th.cat([h, hs.pop()], dim=1), emb, attn_mask, T, frame_indices, attns)
            out = model.out[2](_silu(_gn(model.out[0], h)))
        return out.float().view(B, T, -1, H, W)
    finally:
        th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32 = tf32
