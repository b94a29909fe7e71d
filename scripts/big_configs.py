"""Runs the larger BASELINE configs once (functional check + timing): cfg3 (FFB, 65536 rays), cfg4-like
(StyleModNFFB, T=2^22, 8192 rays/GPU shard) and HashGrid L=16 T=2^19 at 65536 rays."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests_support import make_conf, quiet_build
from tests_support import synthetic_batch


def run(tag, conf, n_rays, steps=3, graph=True):
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, conf).cuda().train()
    tr = DataParallelTrainer(model, IDRLoss(0.1, 100.0, 50.0), lr=1e-4, use_cuda_graph=graph)
    inp, rgb = synthetic_batch(n_rays, seed=1)
    inp = {k: v.cuda() for k, v in inp.items()}
    gt = {"rgb": rgb.cuda()}
    for _ in range(3):
        l = tr.step(inp, gt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        l = tr.step(inp, gt)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    n_par = sum(p.numel() for p in model.parameters())
    print("%-34s rays=%6d  %8.2f ms/step  %9.0f rays/s  loss=%.4f  params=%.1fM  tracer=%s  mem=%.1f GB" % (
        tag, n_rays, dt * 1e3, n_rays / dt, float(l), n_par / 1e6, model.ray_tracer.last_stats,
        torch.cuda.max_memory_allocated() / 2**30), flush=True)


if __name__ == "__main__":
    run("cfg2 HashGrid L6 T2^5", make_conf("HashGrid", 6, 5, 64, 512, 1.0), 2048)
    run("HashGrid L16 T2^19", make_conf("HashGrid", 16, 19, 16, 2048, 1.0), 2048)
    run("HashGrid L16 T2^19", make_conf("HashGrid", 16, 19, 16, 2048, 1.0), 65536)
    run("cfg3 FFB L6 T2^5", make_conf("FFB", 6, 5, 16, 512, 0.45, view_type="FFB"), 2048, graph=True)
    run("cfg3 FFB L6 T2^5", make_conf("FFB", 6, 5, 16, 512, 0.45, view_type="FFB"), 65536, graph=True)
    run("cfg4 StyleModNFFB L6 T2^22 shard", make_conf("StyleModNFFB", 6, 22, 16, 512, 0.45, view_type="StyleModNFFB"), 8192, graph=True)
