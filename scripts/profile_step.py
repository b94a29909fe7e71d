"""Per-entry-point / per-shape device time of one IDR training step (CUDA events), for optimisation work."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _conf():
    """IDRK_PROFILE_CONF=FFB | StyleModNFFB selects a filter-bank configuration instead of the bench's HashGrid one."""
    from tests_support import make_conf
    which = os.environ.get("IDRK_PROFILE_CONF", "")
    if which == "FFB":
        return make_conf("FFB", 6, 5, 16, 512, 0.45, view_type="FFB")
    if which == "StyleModNFFB":
        return make_conf("StyleModNFFB", 6, 22, 16, 512, 0.45, view_type="StyleModNFFB")
    return bench.model_conf()


def main():
    from idrk import kernels as K
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from tests_support import synthetic_batch
    from tests_support import quiet_build
    prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
    K.set_precision(prec)
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, _conf()).cuda().train()
    use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
    tr = DataParallelTrainer(model, IDRLoss(0.1, 100.0, 50.0), lr=1e-4, use_cuda_graph=use_graph)
    inp, rgb = synthetic_batch(int(os.environ.get("IDRK_PROFILE_RAYS", bench.N_RAYS)), seed=1)
    inp = {k: v.cuda() for k, v in inp.items()}
    gt = {"rgb": rgb.cuda()}
    for _ in range(3):
        tr.step(inp, gt)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        tr.step(inp, gt)
    torch.cuda.synchronize()
    print("wall ms/step (plain): %.2f" % ((time.perf_counter() - t0) / 5 * 1e3))
    tr.use_cuda_graph = False
    model.ray_tracer.use_cuda_graph = False          # CUDA events cannot bracket kernels inside a replayed graph
    K.PROFILE.reset(enabled=True, detail=True)
    tr.step(inp, gt)
    prof = K.PROFILE.summary()
    K.PROFILE.reset(enabled=False)
    rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    tot = sum(v["ms"] for v in prof.values())
    print("total kernel ms: %.2f" % tot)
    for k, v in rows[:45]:
        tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else 0
        print("%8.3f ms %5d calls %7.1f us/call %7.1f TF/s  %s" % (v["ms"], v["calls"], v["ms"] / v["calls"] * 1e3, tf, k))


if __name__ == "__main__":
    main()


def sections():
    """Coarse device-time split of a step: trace (sphere graph / sampler+secant / min-sdf) vs shade+bwd vs optimiser."""
    from idrk import kernels as K
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from tests_support import synthetic_batch
    from tests_support import quiet_build
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, bench.model_conf()).cuda().train()
    tr = DataParallelTrainer(model, IDRLoss(0.1, 100.0, 50.0), lr=1e-4, use_cuda_graph=True)
    inp, rgb = synthetic_batch(int(os.environ.get("IDRK_PROFILE_RAYS", bench.N_RAYS)), seed=1)
    inp = {k: v.cuda() for k, v in inp.items()}
    gt = {"rgb": rgb.cuda()}
    for _ in range(4):
        tr.step(inp, gt)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    acc = {"trace": 0.0, "rest": 0.0}
    for _ in range(5):
        a, b, c = ev(), ev(), ev()
        a.record()
        traced = model.trace(inp)
        b.record()
        eik = model._draw_eikonal(bench.N_RAYS, "cuda")
        tr._graphed(traced, eik, gt["rgb"])
        c.record()
        torch.cuda.synchronize()
        acc["trace"] += a.elapsed_time(b) / 5
        acc["rest"] += b.elapsed_time(c) / 5
    print("sections ms:", acc, model.ray_tracer.last_stats)


if __name__ == "__main__" and len(sys.argv) > 3:
    sections()
