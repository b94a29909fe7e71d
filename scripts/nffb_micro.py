"""Fused filter-bank encoder (csrc/nffb.cu) timing and error against the module path: python scripts/nffb_micro.py
(eager timing: small point counts are bound by the host-side wrapper, not the kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idrk import kernels as K  # noqa: E402
from idrk.model.implicit_differentiable_renderer import IDRNetwork  # noqa: E402
from tests_support import make_conf, quiet_build  # noqa: E402

for tag, conf in (("FFB", make_conf("FFB", 6, 5, 16, 512, 0.45, view_type="FFB")),
                  ("StyleModNFFB T=2^22", make_conf("StyleModNFFB", 6, 22, 16, 512, 0.45, view_type="StyleModNFFB"))):
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, conf).cuda()
    ffb = model.implicit_network.embed_model.embedder_obj
    for n in (4096, 32768, 262144):
        x = (torch.rand(n, 3, device="cuda") * 2 - 1) * 0.4
        out = K.nffb_encode_fwd(ffb, x)
        with torch.no_grad():
            ref = ffb(x)
        err = (out[:, :ffb.embeddings_dim] - ref).abs().max().item() / ref.abs().max().item()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            K.nffb_encode_fwd(ffb, x, out=out)
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / 20 * 1e3
        print("%-22s n=%7d  %8.1f us  %7.1f Mpts/s  max err vs module path / max %.2e" % (tag, n, us, n / us, err))
