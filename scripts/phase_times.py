"""Device time of the phases of one graphed IDR training step (bench config), each phase replayed from its OWN CUDA graph
so the launches are as tight as in the product's graphs:

    sphere tracing (init + 10 x (step + line search))  |  sampler sweep  |  secant  |  min-SDF sweep  |  shade + loss + backward  |  clip + Adam

The tail phases are separated by capturing the tail with n_secant_steps = 0 / in eval mode (no min-SDF) and differencing.
    python scripts/phase_times.py [rays]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def capture(fn):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def time_seq(graphs, reps=10):
    """graphs replayed in sequence `reps` times; returns mean ms of each."""
    n = len(graphs)
    acc = [0.0] * n
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i, g in enumerate(graphs):
            g.replay()
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i in range(n):
            acc[i] += evs[i].elapsed_time(evs[i + 1]) / reps
    return acc


def main():
    from idrk import kernels as K
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from idrk.model.ray_tracing import _Evaluator
    from tests_support import quiet_build, synthetic_batch
    rays = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_RAYS
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, bench.model_conf()).cuda().train()
    tr = DataParallelTrainer(model, IDRLoss(0.1, 100.0, 50.0), lr=1e-4, use_cuda_graph=True, sample_seed=1234)
    inp, rgb = synthetic_batch(rays, seed=1)
    inp = {k: v.cuda() for k, v in inp.items()}
    gt = {"rgb": rgb.cuda()}
    for _ in range(4):
        tr.step(inp, gt)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        tr.step(inp, gt)
    e.record()
    torch.cuda.synchronize()
    print("whole step: %.3f ms" % (s.elapsed_time(e) / 10), model.ray_tracer.last_stats)

    rt = model.ray_tracer
    T = next(iter(rt._states.values()))
    ev = _Evaluator(model.implicit_network.sdf)
    assert ev.fast
    model.implicit_network.refresh_inference_weights()
    with torch.no_grad():
        g_sphere = capture(lambda: rt._sphere_trace(T, ev))
        g_tail = capture(lambda: rt._tail_device(T, ev))
        nsec = rt.n_secant_steps
        rt.n_secant_steps = 0
        g_tail_nosec = capture(lambda: rt._tail_device(T, ev))
        rt.n_secant_steps = nsec
        rt.training = False
        g_tail_eval = capture(lambda: rt._tail_device(T, ev))
        rt.training = True
    t = time_seq([g_sphere, g_tail])
    t2 = time_seq([g_sphere, g_tail_nosec])
    t3 = time_seq([g_sphere, g_tail_eval])
    print("sphere tracing graph      : %.3f ms" % t[0])
    print("tail (sampler+secant+min) : %.3f ms" % t[1])
    print("  secant (8 steps)        : %.3f ms" % (t[1] - t2[1]))
    print("  min-SDF sweep           : %.3f ms" % (t[1] - t3[1]))
    print("  sampler sweep + resolve : %.3f ms" % (t[1] - (t[1] - t2[1]) - (t[1] - t3[1])))
    # shade + loss + backward graph of the trainer and the optimiser, in the two-graph form of the step (the whole-step graph
    # cannot be bracketed inside): tracer graph | eager hand-over | shade graph | optimiser
    tr.whole_step_graph = False
    for _ in range(4):
        tr.step(inp, gt)
    torch.cuda.synchronize()
    s.record()
    for _ in range(10):
        tr.step(inp, gt)
    e.record()
    torch.cuda.synchronize()
    print("whole step, two-graph form: %.3f ms" % (s.elapsed_time(e) / 10))
    acc = {"trace_call": 0.0, "shade_bwd_graph": 0.0, "optimiser": 0.0}
    for _ in range(10):
        a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        a.record()
        traced = model.trace(inp)
        b.record()
        eik = model._draw_eikonal(rays, "cuda")
        tr._graphed(traced, eik, gt["rgb"])
        c.record()
        K.sumsq_det(tr.bucket.grad, tr.sumsq, tr.sumsq_partials)
        K.clip_adam(tr.bucket.flat, tr.bucket.grad, tr.m, tr.v, 0.0, 0.9, 0.999, 1e-8, 1, 1.0, tr.sumsq, 1.0)
        d.record()
        torch.cuda.synchronize()
        acc["trace_call"] += a.elapsed_time(b) / 10
        acc["shade_bwd_graph"] += b.elapsed_time(c) / 10
        acc["optimiser"] += c.elapsed_time(d) / 10
    for k, v in acc.items():
        print("%-26s: %.3f ms  (host-synchronised per section: includes launch latency)" % (k, v))


if __name__ == "__main__":
    main()
