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
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        tr.step(inp, gt)
        torch.cuda.synchronize()
    avg = prof.key_averages(group_by_stack_n=12)
    rows = []
    for ev in avg:
        if not ev.key.startswith("aten::") or ev.device_time_total <= 0:
            continue
        site = "?"
        for fr in ev.stack:
            if "hashmodnffbanks-idr_b200" in fr:
                site = fr.split("hashmodnffbanks-idr_b200/")[-1][:90]
                break
        rows.append((ev.count, ev.key, site, ev.device_time_total))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for c, k, sname, t in rows:
        agg[(k, sname)][0] += c
        agg[(k, sname)][1] += t
    print("aten ops with device time, by first frame inside the package (count, device us):")
    for (k, sname), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
        print("%4d %8.1f  %-26s %s" % (c, t, k, sname))


if __name__ == "__main__":
    main()
