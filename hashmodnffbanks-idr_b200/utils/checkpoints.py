"""Checkpoints in the reference's on-disk layout (training/idr_train.py:83-98,181-216 save; :150-178 load):

    <checkpoints_path>/ModelParameters/{epoch}.pth, latest.pth          {"epoch", "model_state_dict"}
    <checkpoints_path>/OptimizerParameters/{epoch}.pth, latest.pth      {"epoch", "optimizer_state_dict"}
    <checkpoints_path>/SchedulerParameters/{epoch}.pth, latest.pth      {"epoch", "scheduler_state_dict"}

`model_state_dict` uses the reference's keys (the modules here keep them, tests/test_dropin_cpu.py), the optimiser
state is `torch.optim.Adam`'s (`DataParallelTrainer.optimizer_state_dict`), so files written here load in the
reference's `IDRTrainRunner` (`--is_continue`) and files written there load here.  Host-side I/O, not on the hot path.
"""
import os
from typing import Dict, Optional

import torch

MODEL_SUBDIR = "ModelParameters"
OPTIMIZER_SUBDIR = "OptimizerParameters"
SCHEDULER_SUBDIR = "SchedulerParameters"


def _save_pair(obj: Dict, root: str, sub: str, epoch) -> None:
    d = os.path.join(root, sub)
    os.makedirs(d, exist_ok=True)
    torch.save(obj, os.path.join(d, "%s.pth" % epoch))
    torch.save(obj, os.path.join(d, "latest.pth"))


def save_checkpoints(checkpoints_path: str, epoch: int, model: torch.nn.Module, optimizer=None, scheduler=None) -> None:
    """`optimizer`: a torch optimiser, a DataParallelTrainer, or a ready state dict; `scheduler`: a torch scheduler,
    a state dict, or None (a MultiStepLR-shaped stub with last_epoch = epoch is written, the reference always
    expects the file)."""
    _save_pair({"epoch": epoch, "model_state_dict": model.state_dict()}, checkpoints_path, MODEL_SUBDIR, epoch)
    if optimizer is not None:
        if hasattr(optimizer, "optimizer_state_dict"):
            osd = optimizer.optimizer_state_dict()
        elif hasattr(optimizer, "state_dict"):
            osd = optimizer.state_dict()
        else:
            osd = optimizer
        _save_pair({"epoch": epoch, "optimizer_state_dict": osd}, checkpoints_path, OPTIMIZER_SUBDIR, epoch)
    if scheduler is None:
        ssd = {"last_epoch": epoch, "_step_count": epoch + 1}
    else:
        ssd = scheduler.state_dict() if hasattr(scheduler, "state_dict") else scheduler
    _save_pair({"epoch": epoch, "scheduler_state_dict": ssd}, checkpoints_path, SCHEDULER_SUBDIR, epoch)


def load_checkpoints(checkpoints_path: str, model: torch.nn.Module, optimizer=None, checkpoint: str = "latest",
                     map_location=None, strict: bool = True) -> int:
    """Loads model (and optimiser) state saved by `save_checkpoints` or by the reference; returns the start epoch
    (idr_train.py:150-178).  After loading into a model that a DataParallelTrainer already re-homed into its flat
    bucket the parameter storage is written in place, so the bucket views stay valid."""
    from .. import mlp
    data = torch.load(os.path.join(checkpoints_path, MODEL_SUBDIR, "%s.pth" % checkpoint), map_location=map_location)
    model.load_state_dict(data["model_state_dict"], strict=strict)
    mlp.weights_changed()                      # folded inference weights must be rebuilt
    opt_file = os.path.join(checkpoints_path, OPTIMIZER_SUBDIR, "%s.pth" % checkpoint)
    if optimizer is not None and os.path.exists(opt_file):
        osd = torch.load(opt_file, map_location=map_location)["optimizer_state_dict"]
        if hasattr(optimizer, "load_optimizer_state_dict"):
            optimizer.load_optimizer_state_dict(osd)
        else:
            optimizer.load_state_dict(osd)
    return int(data["epoch"])
