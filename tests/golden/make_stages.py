"""Stage lists of the reference's sampling schemes, generated from the LIVE reference (sampling_schemes.py:34-230) for the
whole-video driver benchmark / tests: python tests/golden/make_stages.py  (needs /root/reference; writes stages_*.json)."""
import json
import os
import sys
import types

sys.path.insert(0, "/root/reference")
# sampling_schemes imports lpips lazily inside a try (sampling_schemes.py:5-31); nothing else is needed for the fixed schemes
from improved_diffusion.sampling_schemes import sampling_schemes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
for name, T, n_obs, max_frames, step in (("hierarchy-2", 300, 36, 20, 10), ("autoreg", 60, 4, 8, 4)):
    it = iter(sampling_schemes[name](video_length=T, num_obs=n_obs, max_frames=max_frames, step_size=step))
    stages = []
    B = 1
    while True:
        try:
            it.set_videos([None])  # batch of one video (sampling_schemes.py:120-121 only takes its length for the fixed schemes)
            obs, lat = next(it)
        except StopIteration:
            break
        stages.append([list(map(int, obs[0])), list(map(int, lat[0]))])
    done = set(range(n_obs))
    for o, l in stages:
        assert set(o) <= done, "a stage conditions on frames that are not finished yet"
        done |= set(l)
    assert done == set(range(T)), (name, len(done))
    out = dict(scheme=name, video_length=T, n_obs=n_obs, max_frames=max_frames, step_size=step, stages=stages)
    json.dump(out, open(os.path.join(HERE, f"stages_{name}_T{T}.json"), "w"))
    print(name, "stages:", len(stages), "model-frames per diffusion step:", sum(len(o) + len(l) for o, l in stages))
