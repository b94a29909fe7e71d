"""Data-parallel training of the IDR step (north-star subsystem 4; the reference has no multi-GPU code).

One process per GPU.  Parameters are replicated, the ray batch is sharded (each rank traces its own
rays), and after the local backward ONE all-reduce (NCCL over NVLink / NVSwitch) sums a single flat fp32
gradient bucket [hash tables | MLP weights]; the fused clip + Adam kernel (csrc/optim.cu) then applies
the identical update on every rank (1/world folded into the kernel), so replicas stay bit-identical
without broadcasting parameters.  Equal shard sizes are required because IDRLoss normalises by the local
ray count (reference loss.py:19,48): mean over ranks of the local gradients == gradient of the global loss.
"""
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import kernels as K


class FlatBucket:
    """Re-homes a module's parameters (and their .grad) as views of two flat fp32 tensors."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]        # keep every view 16-byte aligned
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        total = self.offsets[-1]
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            view = self.flat[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view_as(p)

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):     # autograd may have replaced .grad; re-attach the views
            g = p.grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view_as(p)

    def gather_stray_grads(self):
        """If autograd installed a fresh .grad tensor (first accumulation) copy it into the bucket."""
        for p, o in zip(self.params, self.offsets):
            g = p.grad
            if g is not None and g.data_ptr() != self.grad.data_ptr() + 4 * o:
                self.grad[o:o + p.numel()].view_as(p).copy_(g)
                p.grad = self.grad[o:o + p.numel()].view_as(p)


def shard_rays(model_input: Dict[str, torch.Tensor], ground_truth: Dict[str, torch.Tensor], rank: int, world: int):
    """Contiguous, equal-sized ray shards of a [B, N, ...] batch (uv / object_mask / rgb)."""
    n = model_input["uv"].shape[1]
    if n % world:
        raise ValueError("ray count %d must be divisible by world size %d (IDRLoss normalises per shard)" % (n, world))
    s = n // world
    sl = slice(rank * s, (rank + 1) * s)
    inp = dict(model_input)
    inp["uv"] = model_input["uv"][:, sl]
    inp["object_mask"] = model_input["object_mask"][:, sl]
    gt = {"rgb": ground_truth["rgb"][:, sl]}
    return inp, gt


def allreduce_mean_(flat_grad: torch.Tensor, world: int, group=None) -> float:
    """Sums the bucket over ranks in place; returns the scale the optimiser must still apply."""
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class DataParallelTrainer:
    """model + loss + clip + Adam, replicated per rank, gradients all-reduced once per step.

    step() = trace (eager, device-driven) -> shade + loss + backward -> all-reduce -> fused clip/Adam.
    With `use_cuda_graph` the shade + loss + backward part - fixed shapes, no host syncs - is captured once
    into a CUDA graph and replayed every step (the launch-bound part of the step: ~600 small kernels)."""

    def __init__(self, model: torch.nn.Module, loss_fn, lr: float = 1e-4, max_norm: float = 1.0, world_size: int = 1,
                 betas=(0.9, 0.999), eps: float = 1e-8, use_cuda_graph: bool = False):
        self.model, self.loss_fn = model, loss_fn
        self.lr, self.max_norm, self.world, self.betas, self.eps = lr, max_norm, world_size, betas, eps
        self.bucket = FlatBucket(list(model.parameters()))
        self.m = torch.zeros_like(self.bucket.flat)
        self.v = torch.zeros_like(self.bucket.flat)
        self.sumsq = torch.zeros(1, device=self.bucket.flat.device, dtype=torch.float32)
        self.t = 0
        self.use_cuda_graph = use_cuda_graph
        if hasattr(model, "ray_tracer"):
            model.ray_tracer.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._static = None
        self._static_out = None
        self._graph_key = None

    # -- optimiser state in torch.optim.Adam's format ---------------------------------------------
    def optimizer_state_dict(self) -> Dict:
        """State of the fused clip + Adam in the layout of `torch.optim.Adam(model.parameters()).state_dict()` - what the
        reference writes to OptimizerParameters/*.pth (training/idr_train.py:185-190) - so a run can move between the
        reference's optimiser and this trainer in either direction."""
        b = self.bucket
        state = {}
        if self.t > 0:
            for i, (p, o) in enumerate(zip(b.params, b.offsets)):
                state[i] = {"step": torch.tensor(float(self.t)),
                            "exp_avg": self.m[o:o + p.numel()].view_as(p).clone(),
                            "exp_avg_sq": self.v[o:o + p.numel()].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(b.params)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: Dict) -> None:
        b = self.bucket
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(b.params):
            raise ValueError("optimizer state has %d parameters, the model %d" % (len(order), len(b.params)))
        self.lr = float(groups[0]["lr"])
        self.betas = tuple(float(x) for x in groups[0]["betas"])
        self.eps = float(groups[0]["eps"])
        self.m.zero_()
        self.v.zero_()
        steps = set()
        for i, (p, o) in zip(order, zip(b.params, b.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("optimizer state %d has shape %s, parameter %s" % (i, tuple(st["exp_avg"].shape), tuple(p.shape)))
            self.m[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
            self.v[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ (%s): the fused optimiser keeps one" % sorted(steps))
        self.t = steps.pop() if steps else 0

    # -- differentiable part ---------------------------------------------------------------------
    def _shade_and_backward(self, traced, eik, rgb):
        K.ZERO_POOL.begin(self.bucket.flat.device)       # one fill for every zero-initialised scratch of the step
        try:
            out = self.model.shade(traced, eik)
            losses = self.loss_fn(out, {"rgb": rgb})
            self.bucket.zero_grad()
            K.DIRECT_GRADS[0] = True                     # kernels accumulate leaf gradients straight into the bucket
            try:
                losses["loss"].backward()
            finally:
                K.DIRECT_GRADS[0] = False
            self.bucket.gather_stray_grads()
        finally:
            K.ZERO_POOL.end()
        return losses

    def _graphed(self, traced, eik, rgb):
        key = tuple((k, tuple(v.shape)) for k, v in sorted(traced.items())) + (tuple(eik.shape), tuple(rgb.shape))
        if self._graph is None or key != self._graph_key:
            self._static = ({k: v.detach().clone() for k, v in traced.items()}, eik.clone(), rgb.clone())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                       # warm-up outside capture (allocator, lazy inits)
                    self._shade_and_backward(*self._static)
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            l0 = K._lib.LAUNCHES[0]
            with torch.cuda.graph(self._graph):
                losses = self._shade_and_backward(*self._static)
                self._static_out = {k: v.detach() for k, v in losses.items()}
            self._graph_launches = K._lib.LAUNCHES[0] - l0
            self._graph_key = key
        st_traced, st_eik, st_rgb = self._static
        for k, v in traced.items():
            st_traced[k].copy_(v)
        st_eik.copy_(eik)
        st_rgb.copy_(rgb)
        self._graph.replay()
        K._lib.LAUNCHES[0] += self._graph_launches          # idrk kernels inside the replayed graph
        return self._static_out

    def step(self, model_input, ground_truth) -> torch.Tensor:
        from . import mlp
        b = self.bucket
        model = self.model
        traced = model.trace(model_input)
        dev = traced["dists"].device
        n = traced["dists"].shape[0]
        eik = model._draw_eikonal(n, dev)
        rgb = ground_truth["rgb"].to(dev)
        if self.use_cuda_graph and model.training:
            losses = self._graphed(traced, eik, rgb)
        else:
            losses = self._shade_and_backward(traced, eik, rgb)
        scale = allreduce_mean_(b.grad, self.world)
        self.t += 1
        self.sumsq.zero_()
        K.sumsq(b.grad, self.sumsq)
        K.clip_adam(b.flat, b.grad, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.t,
                    self.max_norm, self.sumsq, scale)
        mlp.weights_changed()
        self.last_losses = losses
        return losses["loss"]
