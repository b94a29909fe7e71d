"""Which torch ops (not idrk kernels) launch kernels in one training step, and from which line of this package."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from tests_support import synthetic_batch
    from tests_support import quiet_build
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, bench.model_conf()).cuda().train()
    tr = DataParallelTrainer(model, IDRLoss(0.1, 100.0, 50.0), lr=1e-4, use_cuda_graph=False)
    model.ray_tracer.use_cuda_graph = False
    inp, rgb = synthetic_batch(bench.N_RAYS, seed=1)
    inp = {k: v.cuda() for k, v in inp.items()}
    gt = {"rgb": rgb.cuda()}
    for _ in range(3):
        tr.step(inp, gt)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        tr.step(inp, gt)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.key_averages(group_by_input_shape=True):
        if not ev.key.startswith("aten::") or ev.self_device_time_total <= 0:
            continue
        shapes = str([tuple(x) for x in ev.input_shapes if x])[:70]
        agg[(ev.key, shapes)][0] += ev.count
        agg[(ev.key, shapes)][1] += ev.self_device_time_total
    tot = sum(v[1] for v in agg.values())
    print("aten ops with their own device time in one eager step: %.0f us in %d launches-ish" % (tot, sum(v[0] for v in agg.values())))
    print("count   self_us  op  input shapes")
    for (k, shapes), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print("%4d %8.1f  %-28s %s" % (c, t, k, shapes))


if __name__ == "__main__":
    main()
