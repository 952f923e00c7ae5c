"""A short, seeded run of tools/stress_parity.py inside the GPU suite: random alphabets, shapes, symbol
distributions, both coder modes and corrupted streams -- CUDA encoder/decoder against the C oracle."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_tool():
    spec = importlib.util.spec_from_file_location("stress_parity", os.path.join(ROOT, "tools", "stress_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("build", ["lat", "thr"])
def test_random_cases_match_the_oracle(build, monkeypatch):
    from image_compression_2_b200 import codec as _codec
    monkeypatch.setattr(_codec, "DEFAULT_DECODE_FLAGS", {"lat": 1, "thr": 2}[build])  # LC_FLAG_DEC_*_BUILD
    tool = _load_tool()
    rng = np.random.default_rng(20261018 if build == "lat" else 20261019)
    fails, symbols = [], 0
    for _ in range(1200):
        symbols += tool.one_case(rng, fails)
        assert not fails, fails[:3]
    assert symbols > 500000
