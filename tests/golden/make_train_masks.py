"""Generates tests/golden/train_masks.json from the LIVE reference (needs /root/reference; run in the build container):
TrainLoop.sample_all_masks / prepare_training_batch (improved_diffusion/train_util.py:180-245, unmodified) on seeded inputs.
Videos are encoded so that the gathered batch identifies its source: batch1[b, t] = 1000*b + t, batch2[b, t] = -(1000*b + t) - 1.
    python tests/golden/make_train_masks.py
"""
import json
import os
import sys
import types

import numpy as np
import torch as th

comm = types.SimpleNamespace(rank=0, size=1, Get_rank=lambda: 0, Get_size=lambda: 1, bcast=lambda x, root=0: x)
sys.modules["mpi4py"] = types.SimpleNamespace(MPI=types.SimpleNamespace(COMM_WORLD=comm))
sys.modules["mpi4py.MPI"] = sys.modules["mpi4py"].MPI
sys.modules["blobfile"] = types.SimpleNamespace(BlobFile=open, join=os.path.join, dirname=os.path.dirname, exists=os.path.exists)
sys.modules.setdefault("wandb", types.SimpleNamespace(log=lambda *a, **k: None))
sys.path.insert(0, "/root/reference")
from improved_diffusion import train_util  # noqa: E402

assert train_util.__file__.startswith("/root/reference")
CASES = [  # seed, B, T, max_frames, pad_with_random_frames, second batch
    (0, 2, 12, 5, True, True), (1, 4, 20, 5, True, True), (2, 3, 300, 20, True, True), (3, 2, 36, 20, True, False),
    (4, 3, 16, 5, False, False), (5, 2, 41, 40, True, True), (6, 5, 7, 5, True, True), (7, 1, 1000, 20, False, False),
    (8, 4, 30, 10, True, True),
]  # T <= max_frames is left out: upstream's loop cannot terminate once every frame is flagged


def videos(B, T):
    b1 = (1000 * th.arange(B).view(B, 1) + th.arange(T).view(1, T)).float().view(B, T, 1, 1, 1)
    return b1, -b1 - 1


def main():
    out = []
    for seed, B, T, N, pad, second in CASES:
        loop = types.SimpleNamespace(max_frames=N, pad_with_random_frames=pad)
        for name in ("sample_some_indices", "sample_all_masks", "prepare_training_batch"):
            setattr(loop, name, types.MethodType(getattr(train_util.TrainLoop, name), loop))
        b1, b2 = videos(B, T)
        th.manual_seed(seed)
        np.random.seed(seed)
        batch, fi, obs, lat = loop.sample_all_masks(b1, b2 if second else None)
        tail = [float(th.rand(())), float(np.random.rand())]  # RNG positions after the call
        out.append(dict(seed=seed, B=B, T=T, max_frames=N, pad=pad, second=second, batch=batch.flatten(1).long().tolist(),
                        frame_indices=fi.tolist(), obs=obs.flatten(1).long().tolist(), latent=lat.flatten(1).long().tolist(), tail=tail))
    # the index sampler on its own, many draws (exercises the float32 rounding of the truncation)
    th.manual_seed(123)
    np.random.seed(123)
    loop = types.SimpleNamespace()
    loop.sample_some_indices = types.MethodType(train_util.TrainLoop.sample_some_indices, loop)
    draws = [loop.sample_some_indices(n, T) for n, T in [(5, 12), (20, 300), (40, 41), (3, 7), (20, 36)] for _ in range(200)]
    json.dump(dict(cases=out, draws=draws), open(os.path.join(os.path.dirname(__file__), "train_masks.json"), "w"))
    print("wrote", len(out), "cases,", len(draws), "draws")


if __name__ == "__main__":
    main()
