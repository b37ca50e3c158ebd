import gzip
import os
import shutil
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests/` on a box without a CUDA device (or without nvcc to build the library) skips the GPU tests
    instead of erroring in the `device` fixture.  `-m gpu` on a GPU box runs them; there the library must load or the run fails."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:                                      # noqa: BLE001
        have_gpu = False
    lib = os.path.join(REPO, "literate_b200", "_lib", "libliterate_b200.so")
    have_lib = os.path.exists(lib) or shutil.which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "libliterate_b200.so is not built and nvcc is missing"
    skip = pytest.mark.skip(reason="needs a B200: " + why)
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def golden_input(name, tmp_path=None):
    """Path of a reference input table; the large one is stored gzipped and unpacked on demand."""
    p = os.path.join(GOLD, "inputs", name)
    if os.path.exists(p):
        if tmp_path is None:
            return p
        dst = os.path.join(str(tmp_path), name)
        shutil.copy(p, dst)
        return dst
    if tmp_path is None:
        raise FileNotFoundError(name + " is stored gzipped: pass tmp_path")
    dst = os.path.join(str(tmp_path), name)
    with gzip.open(p + ".gz", "rb") as src, open(dst, "wb") as out:
        shutil.copyfileobj(src, out)
    return dst


@pytest.fixture(scope="session")
def device():
    from literate_b200.engine import Device
    return Device(0)


@pytest.fixture(scope="session")
def metal_path(tmp_path_factory):
    return golden_input("metal_bands_1.tsv", tmp_path_factory.mktemp("metal"))


def random_states(rng, n, start, end, kmax=8, rate_scale=1.0):
    """Valid (L, M, timesL, timesM) states: shifts more than 1 apart (LiteRateForward.py:290)."""
    out = []
    for _ in range(n):
        sides = []
        for _s in range(2):
            k = int(rng.integers(1, kmax + 1))
            while True:
                sh = np.sort(rng.uniform(start, end, k - 1))
                t = np.concatenate([[start], sh, [end]])
                if k == 1 or np.min(np.diff(t)) > 1.0:
                    break
                k = max(1, k - 1)
            sides.append((rng.gamma(2.0, rate_scale, k), t))
        out.append((sides[0][0], sides[1][0], sides[0][1], sides[1][1]))
    return out
