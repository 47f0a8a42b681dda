import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MOLFILE = os.path.join(ROOT, "radex_emcee_b200", "data", "co.dat")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _gpu_unavailable():
    """Reason the `gpu` tests cannot run here, or None.  (libradex_b200.so is built for sm_100a only.)"""
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
        if torch.cuda.get_device_capability(0)[0] < 10:
            return "libradex_b200.so is built for sm_100a; device is sm_%d%d" % torch.cuda.get_device_capability(0)
    except Exception as e:      # pragma: no cover
        return "torch unusable: %s" % e
    if not os.path.exists(os.path.join(ROOT, "radex_emcee_b200", "libradex_b200.so")):
        return "libradex_b200.so is not built (python -m radex_emcee_b200.build)"
    return None


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a B200 skips the gpu tests instead of failing them.  With an
    explicit `-m gpu` they run regardless, so a GPU box with a broken build fails loudly instead of skipping."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    why = _gpu_unavailable()
    if why:
        skip = pytest.mark.skip(reason=why)
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def molfile():
    return MOLFILE


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle(MOLFILE)


@pytest.fixture(scope="session")
def golden_solve():
    return np.load(os.path.join(GOLDEN, "macho_solve.npz"))


def draw_params(rng, n, tbg, lo_n=2.0, hi_n=7.0, lo_N=15.5, hi_N=19.5):
    """Config-2 draws."""
    out = np.empty((0, 3))
    while out.shape[0] < n:
        ln = rng.uniform(lo_n, hi_n, 2 * n)
        lt = rng.uniform(np.log10(tbg), 3.0, 2 * n)
        lN = rng.uniform(lo_N, hi_N, 2 * n)
        ok = (lN - ln > 10.0) & (lN - ln < 17.5)
        out = np.vstack([out, np.column_stack([10 ** lt, 10 ** ln, 10 ** lN])[ok]])
    return out[:n]
