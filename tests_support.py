"""Helpers shared by tests/, __graft_entry__.smoke() and bench.py (checker side only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_sd_into(module: torch.nn.Module, sd: dict, prefix: str = "", strict: bool = True):
    """Copies a reference-keyed state dict (optionally under `prefix`) into a product module."""
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    missing, unexpected = module.load_state_dict(sub, strict=False)
    if strict and (missing or unexpected):
        raise AssertionError("state_dict mismatch: missing=%s unexpected=%s" % (missing, unexpected))
    return module


def smoke_check(device):
    """One small invocation of the hot path on `device`, checked against the oracle."""
    from oracle import idr_oracle as O
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    gen = torch.Generator().manual_seed(0)
    L, F, log2T, base, desired = 16, 2, 12, 16, 2048
    sd = O.make_hashgrid_sd("", L, F, log2T, base, desired, gen, table_std=0.5)
    m = MultiResHashGridMLP(True, 3, L, F, log2T, base, desired)
    load_sd_into(m, sd)
    m = m.to(device)
    x = torch.rand(4096, 3, generator=gen) * 2 - 1
    y = m(x.to(device)).cpu()
    ref = O.hashgrid_embed(x, sd, "", L, base, desired)
    assert torch.equal(y[:, 3 + 2 * L:], ref[:, 3 + 2 * L:]), "hash block must be bit exact"
    assert torch.allclose(y[:, :3 + 2 * L], ref[:, :3 + 2 * L], atol=4e-6, rtol=0), "fourier prefix"
    return True
