"""
B200-native drop-in for the hot path of `improved_diffusion` (FDM latent video diffusion).

Only the denoiser + DDPM step math live here (unet, rpe, nn, gaussian_diffusion, respace, script_util),
re-implemented on hand-written sm_100a kernels (libfdm_sm100.so).  Everything else of the reference
package (train_util, sampling_schemes, video_datasets, dist_util, logger, ...) is out of scope and is
used UNCHANGED: point FDM_REFERENCE_PATH at a checkout of the reference and its `improved_diffusion/`
directory is appended to this package's search path, so `from improved_diffusion import train_util`
resolves to the reference file while `improved_diffusion.unet` etc. resolve to this package.
"""
import os as _os

_ref = _os.environ.get("FDM_REFERENCE_PATH")
if _ref:
    _cand = _os.path.join(_ref, "improved_diffusion")
    if _os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
