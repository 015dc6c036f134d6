"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` covers the oracle against the golden vectors, host logic and the C-ABI symbol table;
`-m gpu` are the parity tests proper (CUDA path vs oracle, through the C ABI) and need a B200.
Nothing under `-m gpu` reads /root/reference (it does not exist on the GPU box).
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist before anything imports the package (nvcc cross-compiles on CPU)."""
    from lidar_ai_recommendation_software_b200 import build as _b  # no CUDA needed for this import
    _b.build()


def load_golden(name: str):
    return np.load(GOLDEN / f"{name}.npz", allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


CASES = {
    "ref_sample_10k": lambda synth: synth.reference_sample(),
    "crowd_20k": lambda synth: synth.add_outliers(synth.crowd_frame(20000, seed=3, extent=15.0))[:, :3].astype(np.float64),
    "crowd_100k": lambda synth: synth.crowd_frame(100000, seed=0, extent=50.0)[:, :3].astype(np.float64),
    "tiny_14": lambda synth: synth.tiny_cloud(),
    "sparse_300": lambda synth: synth.sparse_cloud(),
}


@pytest.fixture(scope="session")
def case_points():
    cache = {}

    def get(name):
        if name not in cache:
            from lidar_ai_recommendation_software_b200 import synth
            cache[name] = CASES[name](synth)
        return cache[name]

    return get
