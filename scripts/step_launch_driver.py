"""ncu driver: the kernels of exactly ONE graphed training step (the bench's model, batch and trainer).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/step_launches.csv python scripts/step_launch_driver.py

Warm-up steps (graph capture included) run outside the profiler range; cudaProfilerStart/Stop bracket one step.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from idrk import kernels as K
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from tests_support import synthetic_batch
    from tests_support import quiet_build
    K.set_precision("3xtf32")
    torch.manual_seed(0)
    conf = bench.model_conf()
    which = os.environ.get("IDRK_PROFILE_CONF", "")          # FFB | StyleModNFFB: the filter-bank configurations (cfg3 / cfg4 models)
    if which:
        from tests_support import make_conf
        conf = (make_conf("FFB", 6, 5, 16, 512, 0.45, view_type="FFB") if which == "FFB" else
                make_conf("StyleModNFFB", 6, 22, 16, 512, 0.45, view_type="StyleModNFFB"))
    model = quiet_build(IDRNetwork, conf).cuda().train()
    tr = DataParallelTrainer(model, IDRLoss(0.1, 100.0, 50.0), lr=1e-4, use_cuda_graph=True)
    inp, rgb = synthetic_batch(bench.N_RAYS, seed=1)
    inp = {k: v.cuda() for k, v in inp.items()}
    gt = {"rgb": rgb.cuda()}
    for _ in range(4):
        tr.step(inp, gt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.step(inp, gt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("one step profiled")


if __name__ == "__main__":
    main()
