"""Checkpoint interop and the pixel sampler - host logic, CPU only.
The trainer's step needs CUDA; constructing it and converting its optimiser state does not."""
import os

import numpy as np
import torch

from idrk.dist import DataParallelTrainer
from idrk.utils import checkpoints as ck
from idrk.utils.sampling import DevicePixelSampler, uv_lattice


def _fake_trainer(model):
    """The real trainer object (its constructor only re-homes the parameters into the flat bucket and allocates the
    Adam moments - no kernel is launched), on CPU tensors."""
    return DataParallelTrainer(model, loss_fn=None, lr=1e-4, max_norm=1.0)


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Softplus(), torch.nn.Linear(7, 3))


def test_optimizer_state_roundtrip_with_torch_adam():
    """torch.optim.Adam state -> trainer (m, v, t) -> Adam-format state dict that torch.optim.Adam accepts."""
    model = _model()
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    for _ in range(3):
        opt.zero_grad()
        model(torch.randn(11, 5)).pow(2).sum().backward()
        opt.step()
    t = _fake_trainer(model)
    t.load_optimizer_state_dict(opt.state_dict())
    assert t.t == 3 and abs(t.lr - 3e-4) < 1e-12
    for i, (p, o) in enumerate(zip(t.bucket.params, t.bucket.offsets)):
        assert torch.equal(t.m[o:o + p.numel()].view_as(p), opt.state[p]["exp_avg"])
        assert torch.equal(t.v[o:o + p.numel()].view_as(p), opt.state[p]["exp_avg_sq"])
    sd = t.optimizer_state_dict()
    opt2 = torch.optim.Adam(model.parameters(), lr=1.0)
    opt2.load_state_dict(sd)                                     # torch accepts the layout
    assert opt2.param_groups[0]["lr"] == 3e-4
    for p in model.parameters():
        assert torch.equal(opt2.state[p]["exp_avg"], opt.state[p]["exp_avg"])
        assert float(opt2.state[p]["step"]) == 3.0
    # both continue identically
    g = [torch.randn_like(p) for p in model.parameters()]
    m2 = _model()
    m2.load_state_dict(model.state_dict())
    opt3 = torch.optim.Adam(m2.parameters(), lr=1.0)
    opt3.load_state_dict(sd)
    for p, q, gg in zip(model.parameters(), m2.parameters(), g):
        p.grad, q.grad = gg.clone(), gg.clone()
    opt.step(); opt3.step()
    for p, q in zip(model.parameters(), m2.parameters()):
        assert torch.allclose(p, q, atol=0, rtol=0)


def test_checkpoint_layout_matches_reference(tmp_path):
    model = _model()
    t = _fake_trainer(model)
    t.t = 5
    t.m.normal_(); t.v.uniform_()
    root = str(tmp_path / "checkpoints")
    ck.save_checkpoints(root, 40, model, optimizer=t)
    for sub, key in ((ck.MODEL_SUBDIR, "model_state_dict"), (ck.OPTIMIZER_SUBDIR, "optimizer_state_dict"),
                     (ck.SCHEDULER_SUBDIR, "scheduler_state_dict")):
        for name in ("40.pth", "latest.pth"):
            d = torch.load(os.path.join(root, sub, name))
            assert d["epoch"] == 40 and key in d
    assert (ck.MODEL_SUBDIR, ck.OPTIMIZER_SUBDIR, ck.SCHEDULER_SUBDIR) == \
        ("ModelParameters", "OptimizerParameters", "SchedulerParameters")              # idr_train.py:86-98
    # what the reference's loader does (idr_train.py:150-165): plain load_state_dict on the same keys
    m2 = _model()
    with torch.no_grad():
        for p in m2.parameters():
            p.add_(1.0)
    opt = torch.optim.Adam(m2.parameters(), lr=1.0)
    saved = torch.load(os.path.join(root, "ModelParameters", "latest.pth"))
    m2.load_state_dict(saved["model_state_dict"])
    opt.load_state_dict(torch.load(os.path.join(root, "OptimizerParameters", "latest.pth"))["optimizer_state_dict"])
    for p, q in zip(model.parameters(), m2.parameters()):
        assert torch.equal(p, q)
    # and our loader reads a checkpoint the reference would have written
    root2 = str(tmp_path / "ref_written")
    for sub, obj in (("ModelParameters", {"epoch": 7, "model_state_dict": model.state_dict()}),
                     ("OptimizerParameters", {"epoch": 7, "optimizer_state_dict": opt.state_dict()})):
        os.makedirs(os.path.join(root2, sub))
        torch.save(obj, os.path.join(root2, sub, "latest.pth"))
    m3, t3 = _model(), None
    with torch.no_grad():
        for p in m3.parameters():
            p.zero_()
    t3 = _fake_trainer(m3)
    assert ck.load_checkpoints(root2, m3, optimizer=t3) == 7
    for p, q in zip(model.parameters(), m3.parameters()):
        assert torch.equal(p, q)
    assert t3.t == 5


def test_uv_lattice_and_sampler_match_reference_recipe():
    H, W, V = 6, 9, 3
    uv = uv_lattice((H, W))
    ref = np.mgrid[0:H, 0:W].astype(np.int32)
    ref = torch.from_numpy(np.flip(ref, axis=0).copy()).float().reshape(2, -1).transpose(1, 0)
    assert torch.equal(uv, ref) and uv.shape == (H * W, 2)
    assert uv[1].tolist() == [1.0, 0.0] and uv[W].tolist() == [0.0, 1.0]       # x runs fastest
    gen = torch.Generator().manual_seed(0)
    rgb = torch.rand(V, H * W, 3, generator=gen) * 2 - 1
    masks = torch.rand(V, H * W, generator=gen) > 0.5
    K = torch.eye(4).repeat(V, 1, 1) * torch.arange(1, V + 1).view(V, 1, 1)
    pose = torch.eye(4).repeat(V, 1, 1) + torch.arange(V).view(V, 1, 1)
    s = DevicePixelSampler(rgb, masks, K, pose, (H, W), device="cpu")
    assert len(s) == V and s.total_pixels == H * W
    idx, inp, gt = s.batch([2, 0])
    assert inp["uv"].shape == (2, H * W, 2) and torch.equal(gt["rgb"], rgb[[2, 0]])
    assert torch.equal(inp["object_mask"], masks[[2, 0]]) and torch.equal(inp["intrinsics"], K[[2, 0]])
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    s.change_sampling_idx(20, generator=g1)
    sel = torch.randperm(H * W, generator=g2)[:20]                # the reference's draw (scene_dataset.py:117)
    idx, inp, gt = s.batch([1])
    assert torch.equal(inp["uv"][0], uv[sel]) and torch.equal(gt["rgb"][0], rgb[1][sel])
    assert torch.equal(inp["object_mask"][0], masks[1][sel]) and torch.equal(inp["pose"], pose[[1]])
    s.change_sampling_idx(-1)
    assert s.batch([0])[1]["uv"].shape[1] == H * W
