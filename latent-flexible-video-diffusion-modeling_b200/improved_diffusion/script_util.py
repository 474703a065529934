"""
Factory with the reference's names, keyword arguments and defaults (improved_diffusion/script_util.py:9-90):
`model_and_diffusion_defaults`, `create_model_and_diffusion`, `create_model`, `create_gaussian_diffusion`,
plus the argparse helpers used by scripts/video_train.py and scripts/video_sample.py.
"""
import argparse

from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps
from .unet import UNetVideoModel

# image_size -> channel multipliers per U-Net level (script_util.py:108-117)
_CHANNEL_MULT = {256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4), 32: (1, 2, 2, 2)}

_MODEL_KEYS = ("image_size", "in_channels", "num_channels", "num_res_blocks", "learn_sigma", "class_cond",
               "use_checkpoint", "attention_resolutions", "num_heads", "num_heads_upsample", "use_scale_shift_norm",
               "dropout", "use_rpe_net")
_DIFFUSION_KEYS = ("learn_sigma", "sigma_small", "noise_schedule", "use_kl", "predict_xstart", "rescale_timesteps",
                   "rescale_learned_sigmas", "timestep_respacing", "diffusion_space_kwargs")


def model_and_diffusion_defaults():
    return dict(
        image_size=64, in_channels=3, num_channels=128, num_res_blocks=2, num_heads=4, num_heads_upsample=-1,
        attention_resolutions="16,8", dropout=0.0, learn_sigma=False, sigma_small=False, class_cond=False,
        diffusion_steps=1000,
        diffusion_space_kwargs=dict(diffusion_space=None, pre_encoded=False, pre_encoded_stats_dict=None),
        noise_schedule="linear", timestep_respacing="", use_kl=False, predict_xstart=False, rescale_timesteps=True,
        rescale_learned_sigmas=True, use_checkpoint=False, use_scale_shift_norm=True, use_rpe_net=True,
    )


def create_model_and_diffusion(image_size, class_cond, learn_sigma, sigma_small, in_channels, num_channels,
                               num_res_blocks, num_heads, num_heads_upsample, attention_resolutions, dropout,
                               diffusion_steps, diffusion_space_kwargs, noise_schedule, timestep_respacing, use_kl,
                               predict_xstart, rescale_timesteps, rescale_learned_sigmas, use_checkpoint,
                               use_scale_shift_norm, use_rpe_net):
    given = dict(locals())
    model = create_model(**{k: given[k] for k in _MODEL_KEYS})
    diffusion = create_gaussian_diffusion(steps=diffusion_steps, **{k: given[k] for k in _DIFFUSION_KEYS})
    return model, diffusion


def create_model(image_size, in_channels, num_channels, num_res_blocks, learn_sigma, class_cond, use_checkpoint,
                 attention_resolutions, num_heads, num_heads_upsample, use_scale_shift_norm, dropout, use_rpe_net):
    if image_size not in _CHANNEL_MULT:
        raise ValueError(f"unsupported image size: {image_size}")
    attention_ds = tuple(image_size // int(res) for res in attention_resolutions.split(","))
    return UNetVideoModel(
        in_channels=in_channels, model_channels=num_channels,
        out_channels=in_channels * 2 if learn_sigma else in_channels,
        num_res_blocks=num_res_blocks, attention_resolutions=attention_ds, image_size=image_size, dropout=dropout,
        channel_mult=_CHANNEL_MULT[image_size], use_checkpoint=use_checkpoint, num_heads=num_heads,
        num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm, use_rpe_net=use_rpe_net)


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear", use_kl=False,
                              predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False,
                              timestep_respacing="",
                              diffusion_space_kwargs={"diffusion_space": "pixel", "pre_encoded": False,
                                                      "pre_encoded_stats_dict": None}):
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    else:
        loss_type = gd.LossType.RESCALED_MSE if rescale_learned_sigmas else gd.LossType.MSE
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, timestep_respacing or [steps]),
        betas=gd.get_named_beta_schedule(noise_schedule, steps),
        model_mean_type=gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON,
        model_var_type=var_type, loss_type=loss_type, rescale_timesteps=rescale_timesteps,
        diffusion_space_kwargs=diffusion_space_kwargs)


def add_dict_to_argparser(parser, default_dict):
    for name, default in default_dict.items():
        kind = str if default is None else (str2bool if isinstance(default, bool) else type(default))
        parser.add_argument(f"--{name}", default=default, type=kind)


def args_to_dict(args, keys):
    return {k: getattr(args, k) for k in keys}


def str2bool(v):
    if isinstance(v, bool):
        return v
    low = v.lower()
    if low in ("yes", "true", "t", "y", "1"):
        return True
    if low in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected")
