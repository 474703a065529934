"""
Stage driver of long-video sampling — the caller immediately above the hot path, restating
`sample_video` of scripts/video_sample.py:28-85 with the video buffer RESIDENT ON THE DEVICE.

The reference keeps `samples` on the host: every stage gathers x0 frame by frame on the CPU, uploads it, runs 1000 diffusion
steps, downloads the result (`.cpu()`) and scatters it row by row, with four prints in between.  Once a stage takes ~2 s instead
of minutes those round trips and Python loops are no longer free.  Here:
  * `samples` [B, T, C, H, W] lives on the GPU for the whole video; the per-stage gather (x0 = samples[b, frame_indices[b]])
    and scatter (samples[b, latent_indices[b]] = generated frames) are ONE indexing kernel each;
  * the sampling-scheme iterator is the reference's own object, unchanged (it is host logic; adaptive schemes receive the
    device-resident buffer through `set_videos`, as upstream does with `samples.to(device)`);
  * stages of equal shape reuse the denoiser plan and its CUDA graph (`gaussian_diffusion._graph_sampler`);
  * one device->host copy at the very end.
"""
import torch as th


@th.no_grad()
def sample_video_with_iterator(model, diffusion, batch, frame_indices_iterator, n_obs, *, device=None, clip_denoised=True,
                               just_get_indices=False, progress=False):
    """batch: [B, T, C, H, W] (host or device); the first `n_obs` frames of every video are observed.
    Returns (samples on batch's original device, [(obs_frame_indices, latent_frame_indices), ...]) like the reference."""
    device = th.device(device) if device is not None else next(model.parameters()).device
    B = batch.shape[0]
    src = batch.to(device, non_blocking=True)
    samples = th.zeros_like(src)
    samples[:, :n_obs] = src[:, :n_obs]
    rows = th.arange(B, device=device).unsqueeze(1)
    indices_used = []
    while True:
        frame_indices_iterator.set_videos(samples)  # ignored by the non-adaptive schemes
        try:
            obs_idx, lat_idx = next(frame_indices_iterator)
        except StopIteration:
            break
        obs_t = th.as_tensor(obs_idx, dtype=th.long)
        lat_t = th.as_tensor(lat_idx, dtype=th.long)
        n_lat = lat_t.shape[1]
        frame_indices = th.cat([obs_t, lat_t], dim=1).to(device)
        obs_mask = th.cat([th.ones_like(obs_t), th.zeros_like(lat_t)], dim=1).view(B, -1, 1, 1, 1).float().to(device)
        latent_mask = 1 - obs_mask
        if just_get_indices:
            local = src[rows, frame_indices]
        else:
            x0 = samples[rows, frame_indices]  # one gather kernel; a fresh tensor, as the reference's .clone()
            local, _ = diffusion.p_sample_loop(model, x0.shape, clip_denoised=clip_denoised,
                                               model_kwargs=dict(frame_indices=frame_indices, x0=x0, obs_mask=obs_mask,
                                                                 latent_mask=latent_mask),
                                               latent_mask=latent_mask, return_attn_weights=False, progress=progress)
        samples[rows, frame_indices[:, -n_lat:]] = local[:, -n_lat:]
        indices_used.append((obs_idx, lat_idx))
    return samples.to(batch.device), indices_used


def sample_video(args, model, diffusion, batch, just_get_indices=False):
    """Same call as scripts/video_sample.py::sample_video(args, model, diffusion, batch): builds the reference's sampling-scheme
    iterator from `args` (sampling_scheme, n_obs, max_frames, max_latent_frames, optimality, eval_dir) and runs the
    device-resident driver.  `improved_diffusion.sampling_schemes` is the reference's own module."""
    from improved_diffusion.sampling_schemes import sampling_schemes
    T = batch.shape[1]
    optimal = None if getattr(args, "optimality", None) is None else args.eval_dir / "optimal_schedule.pt"
    it = iter(sampling_schemes[args.sampling_scheme](video_length=T, num_obs=args.n_obs, max_frames=args.max_frames,
                                                      step_size=args.max_latent_frames, optimal_schedule_path=optimal))
    return sample_video_with_iterator(model, diffusion, batch, it, args.n_obs, device=getattr(args, "device", None),
                                      clip_denoised=getattr(args, "clip_denoised", True), just_get_indices=just_get_indices)
