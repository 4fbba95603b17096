"""bench.py's contract, as far as it can be checked without a GPU: both arms describe the same workload, the
reference arm prints a complete JSON line (it runs the oracle port on the host cores: the one place outside tests/
and smoke() where oracle/ may be executed), the exchange label follows the rule compiled into csrc/exchange.cu."""
import argparse
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(gpus=1, steps=20, warmup=5, impl="ours", config=1, views=0, chunk=8, feat=-1, path="fused",
             exchange="factored", transport="auto")
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("config", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("world", [1, 2, 8])
def test_both_arms_print_the_same_config(config, world):
    ours = bench.bench_config(_args(config=config, gpus=world), world)[5]
    ref = bench.bench_config(_args(config=config, gpus=world, impl="reference"), world)[5]
    assert ours == ref
    assert ours["workload"].startswith(f"cfg{config}_")
    assert not any(k in ours for k in ("model", "global_batch", "seq_len"))   # a workload, not a network
    assert ours["channels"] == 7 + bench.bench_config(_args(config=config), 1)[1]
    json.dumps(ours)


def test_config1_is_the_single_view_case_and_config2_shards_its_views():
    cfg, D, V, chunk, strong, config = bench.bench_config(_args(config=1, gpus=8), 8)
    assert (V, chunk, strong) == (1, 1, False) and config["gaussians"] == 500_000 and config["image"] == [640, 480]
    assert "all-gather" in config["gradient_exchange"]
    _, _, V2, chunk2, strong2, config2 = bench.bench_config(_args(config=2, gpus=8), 8)
    assert strong2 and V2 == 8 and chunk2 == 8 and config2["gradient_exchange"] == "none"
    _, _, V3, _, strong3, config3 = bench.bench_config(_args(config=3, gpus=8), 8)
    assert not strong3 and V3 == 8 and config3["gradient_exchange"] == "all-reduce"   # 8 ranks x 8 views: factors no longer pay


def test_exchange_label_follows_the_kernels_rule(monkeypatch):
    monkeypatch.delenv("GG_NVLS_MODE", raising=False)
    mb = 1 << 20
    assert bench._exchange_mode(True, 2, 300 * mb).startswith("peer loads")          # two ranks: always peer loads
    assert bench._exchange_mode(True, 8, 54 * mb).startswith("peer loads")           # config 1's bucket
    assert bench._exchange_mode(True, 8, 300 * mb).startswith("NVLS multimem")       # config 3's bucket
    assert bench._exchange_mode(False, 8, 300 * mb).startswith("peer loads")         # no multicast mapping
    monkeypatch.setenv("GG_NVLS_MODE", "0")
    assert bench._exchange_mode(True, 2, 1 * mb).startswith("NVLS multimem")
    assert bench._exchange_mode(False, 2, 1 * mb).startswith("peer loads")           # cannot be forced without a mapping
    monkeypatch.setenv("GG_NVLS_MODE", "2")
    assert bench._exchange_mode(True, 8, 1 * mb) == "peer loads + multimem.st"


def test_spread_tiles_is_deterministic_and_bounded():
    assert bench.spread_tiles(1200, 5000) is None                 # the whole frame fits
    t = bench.spread_tiles(1200, 48)
    assert t == bench.spread_tiles(1200, 48) and len(t) == 48 and t[0] == 0 and t[-1] < 1200 and t == sorted(set(t))


@pytest.mark.timeout(600)
def test_reference_arm_prints_a_complete_line_on_config0():
    """BASELINE configs[0] is the reference's own CPU-runnable case (50 k Gaussians, forward only): small enough to
    run the arm for real here."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "0",
                        "--steps", "1", "--warmup", "3"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=580)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpixels/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 3 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"] == bench.bench_config(_args(config=0, impl="reference", steps=1, warmup=3), 1)[5]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "nothing extrapolated" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
