"""Quantifies what shading ALL N rays (fixed shapes, CUDA-graph capturable) costs over shading only the N_s surface rays
like the reference does (implicit_differentiable_renderer.py:262-308): summed device time of this library's kernels
(CUDA events around every launch, K.PROFILE) for get_rbg_value forward + backward on N rows vs on N_s rows.

    python scripts/shade_redundancy.py > profiles/r02_shade_all_rays_cost.txt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from idrk import kernels as K
    from idrk import mlp
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from tests_support import quiet_build, synthetic_batch
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, bench.model_conf()).cuda().train()
    inp, _ = synthetic_batch(bench.N_RAYS, seed=1)
    inp = {k: v.cuda() for k, v in inp.items()}
    traced = model.trace(inp)
    surf = traced["network_object_mask"] & traced["object_mask"]
    pts = (traced["cam_loc"].unsqueeze(1) + traced["dists"].reshape(1, -1, 1) * traced["ray_dirs"]).reshape(-1, 3)
    view = -traced["ray_dirs"].reshape(-1, 3)
    n, n_s = pts.shape[0], int(surf.sum())

    def run(p, v):
        for prm in model.parameters():
            prm.grad = None
        with mlp.shared_weights():
            rgb = model.get_rbg_value(p.clone(), v)
        rgb.sum().backward()

    def kernel_ms(p, v):
        for _ in range(3):
            run(p, v)
        torch.cuda.synchronize()
        K.PROFILE.reset(enabled=True)
        for _ in range(5):
            run(p, v)
        torch.cuda.synchronize()
        prof = K.PROFILE.summary()
        K.PROFILE.reset(enabled=False)
        return sum(v_["ms"] for v_ in prof.values()) / 5, sum(v_["calls"] for v_ in prof.values()) // 5
    all_ms, all_calls = kernel_ms(pts, view)
    s_ms, s_calls = kernel_ms(pts[surf].contiguous(), view[surf].contiguous())
    print("get_rbg_value forward + backward (cfg2 model, %d rays, %d surface rays = %.0f %%)" % (n, n_s, 100.0 * n_s / n))
    print("  all N rays     : %.3f ms of idrk kernel time, %d launches" % (all_ms, all_calls))
    print("  N_s rays only  : %.3f ms of idrk kernel time, %d launches" % (s_ms, s_calls))
    print("  cost of fixed shapes: %.3f ms per step (%.0f %% of this branch)" % (all_ms - s_ms, 100.0 * (all_ms - s_ms) / all_ms))


if __name__ == "__main__":
    main()
