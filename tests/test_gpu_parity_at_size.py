"""The fused `render_views` path -- what bench.py times -- against the CPU oracle at the sizes BASELINE.json
names (configs[1..4]), at the north-star tolerances: integers bit-exact, every blended channel <= 1e-4 max-abs
on non-fragile pixels, gradients element by element within 1e-3 relative plus the stated fp32 floor
(tests/at_size.py explains the floor and asserts that it is not what passes the test)."""
import json
import os

import pytest
import torch

from gaussiangrasper_b200 import scenes

import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import at_size  # noqa: E402

pytestmark = pytest.mark.gpu

REPORT = os.environ.get("GG_AT_SIZE_REPORT", "")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from gaussiangrasper_b200 import _lib
    assert _lib.load().gg_check_device() == 0, _lib.load().gg_last_error_string()
    return torch.device("cuda:0")


def _dump(name, rep):
    if REPORT:
        os.makedirs(os.path.dirname(REPORT) or ".", exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps({name: rep}) + "\n")


def test_small_case_same_checks(dev):
    """The at-size machinery on a case small enough to fail fast (3 views, backward)."""
    cams = scenes.orbit_cameras(3, 160, 120, total=7)
    rep = {}
    try:
        at_size.run_case(dev, 20_000, 160, 120, 6, cams, seed=31, backward=True, report=rep)
    finally:
        _dump("small", rep)


def test_config1_render_views_full_size(dev):
    """configs[1]: 500k Gaussians, one 640x480 view, 23 channels, forward + backward -- the benchmarked step."""
    cfg = scenes.CONFIGS[1]
    cams = scenes.orbit_cameras(1, cfg["W"], cfg["H"], total=8)
    rep = {}
    try:
        at_size.run_case(dev, cfg["n"], cfg["W"], cfg["H"], cfg["D"], cams, seed=1235, backward=True, report=rep)
    finally:
        _dump("config1", rep)


def test_config2_chunk_full_size_forward(dev):
    """configs[2]: 2M Gaussians, 1280x720, one 8-view launch group of the 64-view batch, forward only.  Integers of
    all eight views; images of two of them (the oracle's blend of a 1280x720 view costs seconds each)."""
    cfg = scenes.CONFIGS[2]
    cams = scenes.orbit_cameras(8, cfg["W"], cfg["H"], total=64)
    rep = {}
    try:
        at_size.run_case(dev, cfg["n"], cfg["W"], cfg["H"], cfg["D"], cams, seed=1236, backward=False,
                           image_views=(0, 5), report=rep)
    finally:
        _dump("config2", rep)


def test_config3_eight_views_forward_backward(dev):
    """configs[3]: 1M Gaussians, 8 views of 640x480 per GPU, forward + backward, leaf gradients summed over views."""
    cfg = scenes.CONFIGS[3]
    cams = scenes.orbit_cameras(8, cfg["W"], cfg["H"], total=8)
    rep = {}
    try:
        at_size.run_case(dev, cfg["n"], cfg["W"], cfg["H"], cfg["D"], cams, seed=1237, backward=True, report=rep)
    finally:
        _dump("config3", rep)


@pytest.mark.parametrize("D", [3, 32, 64])
def test_config4_channel_sweep_full_size(dev, D):
    """configs[4]: 1M Gaussians, 1920x1080, D = 3 / 32 / 64 feature channels (C = 10 / 39 / 71; 71 > 64 takes the
    column-block launches without hit masks), forward + backward."""
    cfg = scenes.CONFIGS[4]
    cams = scenes.orbit_cameras(1, cfg["W"], cfg["H"], total=8)
    rep = {}
    try:
        at_size.run_case(dev, cfg["n"], cfg["W"], cfg["H"], D, cams, seed=1238, backward=True, report=rep)
    finally:
        _dump(f"config4_D{D}", rep)
