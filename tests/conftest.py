import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the host-side (CPU) tests exercise losses / DDP / the drop-in TrainLoop through the PyTorch-autograd expression of the network;
# the product refuses that path unless asked for (unet.py: no silent fallback)
os.environ.setdefault("FDM_ALLOW_TORCH_TRAIN", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    return load
