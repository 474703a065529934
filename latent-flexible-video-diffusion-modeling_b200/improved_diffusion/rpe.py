"""
Parameter containers for the relative-position-encoding attention (improved_diffusion/rpe.py).

State-dict layout kept: `<attn>.{qkv,proj_out,norm}` and `<attn>.rpe_{q,k,v}.rpe_net.{embed_distances,
embed_diffusion_time,out}` (rpe.py:11-14,47,111-124).  The attention math itself (rpe.py:133-174) runs in
fdm_temporal_gn / fdm_conv (qkv, proj) / fdm_attn_temporal / fdm_attn_spatial — see engine.py.
"""
import torch.nn as nn

from .nn import normalization, zero_module


class RPENet(nn.Module):
    def __init__(self, channels, num_heads, time_embed_dim):
        super().__init__()
        self.channels, self.num_heads = channels, num_heads
        self.embed_distances = nn.Linear(3, channels)
        self.embed_diffusion_time = nn.Linear(time_embed_dim, channels)
        self.silu = nn.SiLU()
        self.out = zero_module(nn.Linear(channels, channels))  # rpe.py:15-16


class RPE(nn.Module):
    def __init__(self, channels, num_heads, time_embed_dim, use_rpe_net=False):
        super().__init__()
        self.num_heads, self.head_dim, self.use_rpe_net = num_heads, channels // num_heads, use_rpe_net
        if not use_rpe_net:
            # the reference's lookup-table branch reads an attribute that is never set (rpe.py:50); same failure here
            raise AttributeError("'RPE' object has no attribute 'beta'")
        self.rpe_net = RPENet(channels, num_heads, time_embed_dim)


class RPEAttention(nn.Module):
    def __init__(self, channels, num_heads, use_checkpoint=False, time_embed_dim=None, use_rpe_net=None,
                 use_rpe_q=True, use_rpe_k=True, use_rpe_v=True):
        super().__init__()
        self.channels, self.num_heads = channels, num_heads
        self.scale = (channels // num_heads) ** -0.5
        self.use_checkpoint = use_checkpoint
        self.qkv = nn.Linear(channels, channels * 3)
        self.proj_out = zero_module(nn.Linear(channels, channels))
        self.norm = normalization(channels)
        wants_rpe = use_rpe_q or use_rpe_k or use_rpe_v
        if wants_rpe:
            assert use_rpe_net is not None
        mk = lambda on: RPE(channels, num_heads, time_embed_dim, use_rpe_net) if on else None
        self.rpe_q, self.rpe_k, self.rpe_v = mk(use_rpe_q), mk(use_rpe_k), mk(use_rpe_v)

    @property
    def has_rpe(self):
        return self.rpe_q is not None
