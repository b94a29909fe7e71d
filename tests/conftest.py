import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def sd_from(gold, prefix):
    """state_dict stored under 'prefix/<key>' -> dict of torch tensors."""
    out = {}
    for k, v in gold.items():
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(np.array(v))
    return out


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def assert_flips_borderline(mask, mask_ref, margin, tol, what=""):
    """Hit/miss masks may differ from the oracle's only on rays whose decisive SDF values came within `tol` of a decision
    boundary in the oracle's own trace (RayTracerOracle.margin): count AND cause.  Returns the number of flips."""
    mask, mask_ref = mask.reshape(-1).cpu(), mask_ref.reshape(-1).cpu()
    flipped = mask != mask_ref
    n = int(flipped.sum())
    if n:
        worst = float(margin[flipped].max())
        assert worst <= tol, "%s: %d flipped rays, one of them %.3g away from every decision boundary (tol %.3g)" % (
            what, n, worst, tol)
    return n
