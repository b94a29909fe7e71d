"""Importable alias of the product package.

The product lives in ``hashmodnffbanks-idr_b200/`` (a directory name Python cannot import
directly); this shim makes it available as ``idrk`` so that the reference's config strings work
unchanged apart from the prefix, e.g.

    train.model_class = idrk.model.implicit_differentiable_renderer.IDRNetwork
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hashmodnffbanks-idr_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
