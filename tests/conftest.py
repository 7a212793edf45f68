import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ncn():
    import ncn_b200
    return ncn_b200


@pytest.fixture(scope="session")
def vren_ref():
    """The reference's own csrc kernels compiled for sm_100 (oracle/_ref/vren_ref.so) - checker only."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/vren_ref.so not built (run oracle/build_ref.py where /root/reference exists)")
    return mod


@pytest.fixture(scope="session")
def scene():
    """Synthetic room: bitfield + helpers, on cuda:0."""
    import torch
    import ncn_b200
    from ncn_b200 import synth
    occ = synth.room_occupancy(128, 0.5, seed=0)
    grid = synth.density_grid_from_occupancy(occ)
    bits = synth.packbits_np(grid, 5.9)
    dev = torch.device("cuda:0")
    return dict(occ=occ, grid=torch.from_numpy(grid).to(dev), bitfield=torch.from_numpy(bits).to(dev),
                center=torch.zeros(1, 3, device=dev), half_size=torch.full((1, 3), 0.5, device=dev),
                scale=0.5, grid_size=128, cascades=1, max_samples=1024, dev=dev)
