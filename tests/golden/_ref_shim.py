"""Import shim for the *real* reference (only usable where /root/reference exists).

Used by tests/golden/make_golden.py to generate the committed golden vectors and by
the optional `test_oracle_vs_live_reference` tests.  Nothing on the GPU box imports it.

Recipe follows SURVEY.md §8(c): stub the absent third-party modules, make `.cuda()`
a no-op on a CPU-only box and silence the constructor / ray-tracer prints.
"""
import builtins
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("IDRK_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_CODE, "model"))


class DictConf(dict):
    """Stand-in for a pyhocon ConfigTree (only the getters IDRNetwork uses)."""

    def _walk(self, key):
        node = self
        for part in key.split("."):
            node = node[part]
        return node

    def get_int(self, key):
        return int(self._walk(key))

    def get_float(self, key):
        return float(self._walk(key))

    def get_string(self, key):
        return str(self._walk(key))

    def get_list(self, key):
        return list(self._walk(key))

    def get_config(self, key):
        try:
            node = self._walk(key)
        except KeyError:
            return None
        return DictConf(node) if isinstance(node, dict) else node


def load():
    """Returns a namespace with the reference modules (imported once)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch

    for name in ("tinycudann", "imageio", "skimage"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    with contextlib.redirect_stdout(io.StringIO()):
        import model.embeddings.hashGridEmbedding as hge
        import model.embeddings.frequency_enc as fe
        import model.embeddings.nffb3d as nffb
        import model.embeddings.style_Attention.styleMod as sm
        import model.custom_embedder_decoder as ced
        import model.implicit_differentiable_renderer as idr
        import model.ray_tracing as rt
        import model.sample_network as sn
        import model.loss as loss
        import model.density_net as dn
        import utils.rend_util as ru
    ns = types.SimpleNamespace(hge=hge, fe=fe, nffb=nffb, sm=sm, ced=ced, idr=idr, rt=rt,
                               sn=sn, loss=loss, dn=dn, ru=ru, DictConf=DictConf)
    return ns


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
