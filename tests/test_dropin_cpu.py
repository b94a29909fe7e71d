"""CPU checks of the drop-in boundary: config-string resolution, constructor signatures, state-dict keys."""
import numpy as np
import torch

from tests_support import make_conf, quiet_build


def test_drop_in_config_string_and_state_dict_keys(golden):
    """The reference resolves `train.model_class` by dotted name and checkpoints by state_dict keys."""
    from idrk.utils.general import get_class
    g = golden("idr_step")
    cls = get_class("idrk.model.implicit_differentiable_renderer.IDRNetwork")
    model = quiet_build(cls, make_conf("HashGrid", 6, 5, 64, 512, 1.0, width=96, feature=32))
    ref_keys = sorted(k[len("sd_hash/"):] for k in g if k.startswith("sd_hash/"))
    assert sorted(model.state_dict().keys()) == ref_keys
    model = quiet_build(cls, make_conf("StyleModNFFB", 6, 5, 16, 512, 0.45, width=96, feature=32))
    ref_keys = sorted(k[len("sd_style/"):] for k in g if k.startswith("sd_style/"))
    assert sorted(model.state_dict().keys()) == ref_keys
