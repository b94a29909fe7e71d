"""Stability check of the graphed trainer: a few hundred steps of the bench config with a fresh batch of pixels every step
(the reference loop's access pattern); losses must stay finite and go down.  python scripts/long_run_check.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from idrk.dist import DataParallelTrainer  # noqa: E402
from idrk.model.implicit_differentiable_renderer import IDRNetwork  # noqa: E402
from idrk.model.loss import IDRLoss  # noqa: E402
from tests_support import quiet_build, synthetic_batch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
torch.manual_seed(0)
model = quiet_build(IDRNetwork, bench.model_conf()).cuda().train()
tr = DataParallelTrainer(model, IDRLoss(0.1, 100.0, 50.0), lr=1e-4, use_cuda_graph=True, sample_seed=1)
batches = [synthetic_batch(bench.N_RAYS, seed=100 + i) for i in range(8)]
batches = [({k: v.cuda() for k, v in b[0].items()}, {"rgb": b[1].cuda()}) for b in batches]
losses = []
for i in range(steps):
    inp, gt = batches[i % len(batches)]
    losses.append(tr.step(inp, gt))
    if i == steps // 2:
        tr.loss_fn.alpha *= 2.0            # the reference doubles alpha at its milestones: the graph must be re-captured
losses = torch.stack([l.detach() for l in losses]).cpu()
assert torch.isfinite(losses).all(), "non-finite loss"
first, last = losses[:20].mean().item(), losses[-20:].mean().item()
print("steps %d  loss first20 %.4f  last20 %.4f  params finite %s  tracer %s" % (
    steps, first, last, bool(torch.isfinite(tr.bucket.flat).all()), model.ray_tracer.last_stats))
assert last < first
