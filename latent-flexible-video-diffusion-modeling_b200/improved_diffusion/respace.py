"""
SpacedDiffusion — a diffusion process over a subset of the base timesteps (improved_diffusion/respace.py).

Same API (space_timesteps, SpacedDiffusion, _WrappedModel, timestep_map).  The timestep map lookup that the
reference performs with a fresh `th.tensor(timestep_map)` H2D copy on every model call (respace.py:118-124) is a
device-resident table here; in the CUDA-graph sampler the lookup happens inside fdm_timestep_embedding.
"""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Which base steps to keep: per-section even striding, or "ddimN" fixed striding (respace.py:7-60)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                picked = range(0, num_timesteps, stride)
                if len(picked) == want:
                    return set(picked)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(tok) for tok in section_counts.split(",")]
    n_sec = len(section_counts)
    base, rem = divmod(num_timesteps, n_sec)
    kept, offset = [], 0
    for k, count in enumerate(section_counts):
        length = base + (1 if k < rem else 0)
        if length < count:
            raise ValueError(f"cannot divide section of {length} steps into {count}")
        step = 1 if count <= 1 else (length - 1) / (count - 1)
        kept.extend(offset + round(pos) for pos in _accumulate(step, count))
        offset += length
    return set(kept)


def _accumulate(step, count):
    """0, step, 2*step ... by repeated addition — the reference accumulates (cur_idx += frac_stride), and
    round(accumulated) can differ from round(j*step) in the last bit, so the same summation order is kept."""
    pos, out = 0.0, []
    for _ in range(count):
        out.append(pos)
        pos += step
    return out


class SpacedDiffusion(GaussianDiffusion):
    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base_acp = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64), axis=0)
        self.timestep_map, new_betas, prev = [], [], 1.0
        for i, acp in enumerate(base_acp):
            if i in self.use_timesteps:
                new_betas.append(1 - acp / prev)
                prev = acp
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _model_t_host(self):
        m = th.tensor(self.timestep_map)
        return m.float() * (1000.0 / self.original_num_steps) if self.rescale_timesteps else m.float()

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def p_sample(self, model, *args, **kwargs):
        return super().p_sample(self._wrap_model(model), *args, **kwargs)

    def training_losses(self, model, *args, **kwargs):
        return super().training_losses(self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t  # done by the wrapped model


class _WrappedModel:
    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model, self.timestep_map = model, timestep_map
        self.rescale_timesteps, self.original_num_steps = rescale_timesteps, original_num_steps
        self._maps = {}

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x, timesteps, **kwargs):
        key = (str(timesteps.device), timesteps.dtype)
        if key not in self._maps:
            self._maps[key] = th.tensor(self.timestep_map, device=timesteps.device, dtype=timesteps.dtype)
        new_ts = self._maps[key][timesteps]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, timesteps=new_ts, **kwargs)
