"""world_size-2 gloo tests (CPU) of the data-parallel host logic: ray sharding, flat gradient bucket,
all-reduce averaging.  DP(2 ranks) gradients on half batches == single-process gradients on the full batch."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from idrk.dist import FlatBucket, allreduce_mean_, shard_rays


def toy_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))


def toy_loss(model, uv, rgb):
    # normalised by the LOCAL number of rays, like IDRLoss (reference loss.py:19,48)
    pred = model(torch.cat([uv, uv ** 2, uv[..., :1]], -1))
    return (pred - rgb).abs().sum() / float(uv.shape[1])


def batch(n=64):
    g = torch.Generator().manual_seed(3)
    return ({"uv": torch.rand(1, n, 2, generator=g), "object_mask": torch.rand(1, n, generator=g) > 0.5,
             "pose": torch.eye(4)[None], "intrinsics": torch.eye(4)[None]},
            {"rgb": torch.rand(1, n, 3, generator=g)})


def _worker(rank, world, init_file, ret):
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    model = toy_model()
    bucket = FlatBucket(list(model.parameters()))
    inp, gt = batch()
    sinp, sgt = shard_rays(inp, gt, rank, world)
    assert sinp["uv"].shape[1] == 32 and sinp["object_mask"].shape[1] == 32
    bucket.zero_grad()
    toy_loss(model, sinp["uv"], sgt["rgb"]).backward()
    bucket.gather_stray_grads()
    scale = allreduce_mean_(bucket.grad, world)
    ret[rank] = (bucket.grad * scale).clone()
    dist.destroy_process_group()


def test_dp_gradients_equal_single_process():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "init"), ret), nprocs=world, join=True)
        g0, g1 = ret[0], ret[1]
    assert torch.equal(g0, g1)                        # replicas see the same reduced bucket
    model = toy_model()
    bucket = FlatBucket(list(model.parameters()))
    inp, gt = batch()
    bucket.zero_grad()
    toy_loss(model, inp["uv"], gt["rgb"]).backward()
    bucket.gather_stray_grads()
    assert torch.allclose(g0, bucket.grad, atol=1e-6, rtol=1e-5)


def test_flat_bucket_views_and_alignment():
    model = toy_model()
    ref = [p.detach().clone() for p in model.parameters()]
    b = FlatBucket(list(model.parameters()))
    for p, r, o in zip(model.parameters(), ref, b.offsets):
        assert torch.equal(p.detach(), r)
        assert p.data_ptr() == b.flat.data_ptr() + 4 * o and o % 4 == 0
    b.flat.mul_(2.0)
    for p, r in zip(model.parameters(), ref):
        assert torch.equal(p.detach(), 2 * r)


def test_shard_rays_rejects_uneven():
    inp, gt = batch(63)
    try:
        shard_rays(inp, gt, 0, 2)
    except ValueError:
        return
    raise AssertionError("uneven shards must be rejected")


def test_multistep_lr_shim_matches_torch_scheduler():
    """idrk.dist.MultiStepLR drives `trainer.lr` exactly like torch's MultiStepLR drives Adam's lr in the reference loop
    (training/idr_train.py:131-134,312), and its state dict loads into torch's scheduler and back."""
    from types import SimpleNamespace
    from idrk.dist import MultiStepLR
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-4)
    ref = torch.optim.lr_scheduler.MultiStepLR(opt, [3, 5, 5, 9], gamma=0.5)
    tr = SimpleNamespace(lr=1e-4)
    mine = MultiStepLR(tr, [3, 5, 5, 9], gamma=0.5)
    for epoch in range(12):
        assert abs(tr.lr - opt.param_groups[0]["lr"]) <= 1e-12 * 1e-4, epoch
        opt.step()
        ref.step()
        mine.step()
    sd = mine.state_dict()
    opt2 = torch.optim.Adam([p], lr=1e-4)
    ref2 = torch.optim.lr_scheduler.MultiStepLR(opt2, [1], gamma=0.1)
    ref2.load_state_dict(sd)
    assert ref2.last_epoch == mine.last_epoch and ref2.gamma == 0.5 and dict(ref2.milestones) == {3: 1, 5: 2, 9: 1}
    tr3 = SimpleNamespace(lr=1e-4)
    mine3 = MultiStepLR(tr3, [1], gamma=0.1)
    mine3.load_state_dict(ref.state_dict())
    assert mine3.last_epoch == ref.last_epoch and abs(tr3.lr - opt.param_groups[0]["lr"]) <= 1e-16


def _trainer_worker(rank, world, init_file, ret):
    from idrk.dist import DataParallelTrainer
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                        # replicas start DIFFERENT on purpose
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    tr = DataParallelTrainer(model, loss_fn=None, lr=1e-3, world_size=world, sample_seed=5)
    draw = torch.empty(4).uniform_(0, 1, generator=tr.sample_generator)
    ret[rank] = (tr.bucket.flat.clone(), draw, tr.parameters_checksum())
    dist.destroy_process_group()


def test_trainer_broadcasts_parameters_and_seeds_ranks_independently():
    """DataParallelTrainer at world 2 (gloo): rank 0's parameters are broadcast at construction (replicas must not rely on
    identical seeding), the float64 checksum used by bench.py's `replicas_identical` agrees, and every rank owns its own
    generator for the host draws (eikonal points, min-SDF steps) so shards see independent samples."""
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_trainer_worker, args=(world, os.path.join(d, "init"), ret), nprocs=world, join=True)
        (f0, d0, c0), (f1, d1, c1) = ret[0], ret[1]
    assert torch.equal(f0, f1) and torch.equal(c0, c1)
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    assert torch.equal(f0[:35], ref[0].weight.detach().reshape(-1))
    assert not torch.equal(d0, d1)
