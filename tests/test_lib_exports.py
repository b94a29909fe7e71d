"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol that
include/idrk.h declares.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "idrk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(idrk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "idrk_hash_encode_fwd" in syms and "idrk_hash_encode_bwd" in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from idrk import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), "libidrk.so does not export %s" % s
    assert lib.idrk_version() == 4
    assert sorted(_lib.EXPORTS) == header_symbols()


def test_no_fallback_without_cuda():
    """The product path must fail loudly on CPU tensors instead of computing something."""
    import torch
    from idrk._lib import IdrkError
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    m = MultiResHashGridMLP(True, 3, 4, 2, 5, 16, 64)
    with pytest.raises(IdrkError):
        m(torch.rand(8, 3))


def test_product_never_imports_oracle():
    bad = []
    pkg = os.path.join(ROOT, "hashmodnffbanks-idr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_oracle_is_imported_only_by_the_checker_side():
    """Outside tests/, only __graft_entry__/tests_support (smoke) and bench.py's cpu_baseline / reference-arm functions
    may import the oracle; scripts/ (profiling drivers) take their synthetic inputs from tests_support instead."""
    import ast
    for dp, _, files in os.walk(os.path.join(ROOT, "scripts")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"oracle_cfg", "cpu_step_fn", "run_reference"}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            for sub in ast.walk(node):
                if isinstance(sub, ast.ImportFrom) and (sub.module or "").split(".")[0] == "oracle":
                    assert node.name in allowed, node.name
    for node in tree.body:                                   # no module-level import either
        assert not (isinstance(node, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(node))


def test_synthetic_batch_recipe_is_shared():
    import torch
    from oracle import idr_oracle as O
    from tests_support import synthetic_batch
    a, ra = synthetic_batch(300, seed=4)
    b, rb = O.synthetic_batch(300, seed=4)
    assert torch.equal(ra, rb) and all(torch.equal(a[k], b[k]) for k in a)
