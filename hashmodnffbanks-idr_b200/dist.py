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


class MultiStepLR:
    """The reference's learning-rate schedule (training/idr_train.py:131-134: torch.optim.lr_scheduler.MultiStepLR on
    the Adam optimiser, stepped once per epoch at :312) for the fused optimiser: multiplies `trainer.lr` by `gamma` at
    every milestone.  `state_dict()` carries torch's keys so SchedulerParameters/*.pth files move both ways."""

    def __init__(self, trainer, milestones, gamma: float = 0.1, last_epoch: int = -1):
        self.trainer, self.gamma = trainer, float(gamma)
        self.milestones = sorted(int(m) for m in milestones)
        self.base_lrs = [trainer.lr]
        self.last_epoch = last_epoch
        self.step()

    def _lr_at(self, epoch: int) -> float:
        return self.base_lrs[0] * self.gamma ** sum(1 for m in self.milestones if m <= epoch)

    def step(self) -> None:
        self.last_epoch += 1
        self.trainer.lr = self._lr_at(self.last_epoch)

    def get_last_lr(self):
        return [self.trainer.lr]

    def state_dict(self) -> Dict:
        from collections import Counter
        return {"milestones": Counter(self.milestones), "gamma": self.gamma, "base_lrs": list(self.base_lrs),
                "last_epoch": self.last_epoch, "_step_count": self.last_epoch + 1, "_last_lr": [self.trainer.lr]}

    def load_state_dict(self, sd: Dict) -> None:
        self.milestones = sorted(int(m) for m, c in dict(sd["milestones"]).items() for _ in range(int(c)))
        self.gamma = float(sd["gamma"])
        self.base_lrs = [float(x) for x in sd["base_lrs"]]
        self.last_epoch = int(sd["last_epoch"])
        self.trainer.lr = self._lr_at(self.last_epoch)


class DataParallelTrainer:
    """model + loss + clip + Adam, replicated per rank, gradients all-reduced once per step.

    step() = trace (eager, device-driven) -> shade + loss + backward -> all-reduce -> fused clip/Adam.
    With `use_cuda_graph` the shade + loss + backward part - fixed shapes, no host syncs - is captured once
    into a CUDA graph and replayed every step (the launch-bound part of the step: ~600 small kernels)."""

    def __init__(self, model: torch.nn.Module, loss_fn, lr: float = 1e-4, max_norm: float = 1.0, world_size: int = 1,
                 betas=(0.9, 0.999), eps: float = 1e-8, use_cuda_graph: bool = False, rank: Optional[int] = None,
                 sample_seed: Optional[int] = None, broadcast_parameters: bool = True):
        self.model, self.loss_fn = model, loss_fn
        # `lr` is read at every step() (the optimiser kernel takes it by value, outside any CUDA graph): assign
        # `trainer.lr` - or drive it with `MultiStepLR(trainer, ...)` below - like the reference's scheduler.step()
        # (training/idr_train.py:131-134,312).  `param_groups` mirrors torch.optim's attribute for code that
        # rescales `group["lr"]` in place.
        self.param_groups = [{"lr": float(lr)}]
        self.max_norm, self.world, self.betas, self.eps = max_norm, world_size, betas, eps
        self.rank = (dist.get_rank() if (world_size > 1 and dist.is_initialized()) else 0) if rank is None else rank
        self.bucket = FlatBucket(list(model.parameters()))
        if world_size > 1 and broadcast_parameters and dist.is_initialized():
            # replicas must start identical; do not rely on identical seeding / checkpoint loads
            dist.broadcast(self.bucket.flat, src=0)
        # Host-side draws of the step (eikonal points :279, min-SDF steps ray_tracing.py:277).  With `sample_seed`
        # every rank owns a generator seeded `sample_seed + rank`, so data-parallel shards see INDEPENDENT eikonal
        # samples; with None the global host generator is used exactly like the reference (world 1 parity runs).
        self.sample_generator = None
        if sample_seed is not None or world_size > 1:
            self.sample_generator = torch.Generator().manual_seed((0 if sample_seed is None else sample_seed) + 7919 * self.rank)
        self.m = torch.zeros_like(self.bucket.flat)
        self.v = torch.zeros_like(self.bucket.flat)
        self.sumsq = torch.zeros(1, device=self.bucket.flat.device, dtype=torch.float32)
        # the squared gradient norm is summed in a FIXED order (per-CTA partials + one block): every replica derives the
        # same clip factor from the all-reduced bucket, so parameters stay bit-identical across ranks
        self.sumsq_partials = torch.zeros(1184, device=self.bucket.flat.device, dtype=torch.float32)
        self.t = 0
        self.use_cuda_graph = use_cuda_graph
        if hasattr(model, "ray_tracer"):
            model.ray_tracer.use_cuda_graph = use_cuda_graph
            if self.sample_generator is not None:
                model.sample_generator = model.ray_tracer.sample_generator = self.sample_generator
        self._graph = None
        self._static = None
        self._static_out = None
        self._graph_key = None
        self._retired_graphs = []
        # ONE graph for trace + shade + loss + backward (fixed cameras, device-driven tracer); IDRK_WHOLE_GRAPH=0 keeps the
        # two-graph form (tracer graph, eager glue, shade graph) for A/B runs
        import os
        self.whole_step_graph = os.environ.get("IDRK_WHOLE_GRAPH", "1") != "0"
        self._wgraph = None
        self._wkey = None
        self._wstatic = None
        self._wout = None
        self._wlaunches = 0

    @property
    def lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    @lr.setter
    def lr(self, value: float) -> None:
        self.param_groups[0]["lr"] = float(value)

    def parameters_checksum(self) -> torch.Tensor:
        """float64 [sum, sum of squares] of the flat parameter bucket (replica-consistency checks)."""
        f = self.bucket.flat.double()
        return torch.stack([f.sum(), (f * f).sum()])

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
            from . import mlp
            mlp.LEAF_SIDE.begin(self.bucket.flat.device)  # weight / bias gradient work runs on a side stream (mlp._LeafSide)
            try:
                losses["loss"].backward()
                mlp.LEAF_SIDE.finish()
            finally:
                mlp.LEAF_SIDE.active = False
                K.DIRECT_GRADS[0] = False
            self.bucket.gather_stray_grads()
        finally:
            K.ZERO_POOL.end()
        return losses

    def _graphed(self, traced, eik, rgb):
        # Everything the captured launch sequence bakes in as a host scalar is part of the key: the loss weights and
        # alpha (the reference doubles alpha at every alpha_milestone, training/idr_train.py:227-228), the mode of the
        # model and the generation of the shared scratch buffers whose addresses the graph holds.
        lf = self.loss_fn
        key = tuple((k, tuple(v.shape)) for k, v in sorted(traced.items())) + (tuple(eik.shape), tuple(rgb.shape)) + (
            float(getattr(lf, "alpha", 0.0)), float(getattr(lf, "eikonal_weight", 0.0)), float(getattr(lf, "mask_weight", 0.0)),
            bool(self.model.training), K.get_precision(), K.SCRATCH_GENERATION[0])
        if self._graph is None or key != self._graph_key:
            if self._graph is not None:
                self._retired_graphs.append((self._graph, self._static))      # keep its pool alive; never replayed again
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
            self._graph_key = key[:-1] + (K.SCRATCH_GENERATION[0],)     # the warm-up passes may have grown the arena
            K.note_graph_captured()
        st_traced, st_eik, st_rgb = self._static
        for k, v in traced.items():
            st_traced[k].copy_(v)
        st_eik.copy_(eik)
        st_rgb.copy_(rgb)
        self._graph.replay()
        K._lib.LAUNCHES[0] += self._graph_launches          # idrk kernels inside the replayed graph
        return self._static_out

    # -- whole step in one CUDA graph ---------------------------------------------------------------
    def _whole_step_supported(self, model_input) -> bool:
        model = self.model
        if not (self.whole_step_graph and self.use_cuda_graph and model.training and hasattr(model, "ray_tracer")):
            return False
        uv, pose, intr = model_input["uv"], model_input["pose"], model_input["intrinsics"]
        if not (uv.is_cuda and uv.dtype == torch.float32 and pose.dtype == torch.float32 and intr.dtype == torch.float32):
            return False
        if pose.requires_grad or uv.requires_grad or intr.requires_grad:       # trainable cameras keep the eager ray set-up
            return False
        from .model.ray_tracing import _Evaluator
        return _Evaluator(model.implicit_network.sdf).fast

    def _trace_and_backward(self, inp, eik, rgb):
        model = self.model
        model.implicit_network.refresh_inference_weights(force=True)    # weight folding of the tracer's SDF pipeline
        traced = model.trace(inp)
        return self._shade_and_backward(traced, eik, rgb)

    def _graphed_step(self, model_input, rgb):
        """trace + shade + loss + backward replayed from ONE graph.  Host randomness (min-SDF steps, ray_tracing.py:277;
        eikonal points, implicit_differentiable_renderer.py:279) is drawn per step in the reference's order and copied
        into static device buffers; everything else of the step lives in the graph: camera rays, the tracer's launch
        sequence, the hand-over of its results to the differentiable part (no clones / copies between two graphs)."""
        model, rt, lf = self.model, self.model.ray_tracer, self.loss_fn
        dev = self.bucket.flat.device
        names = ("uv", "pose", "intrinsics", "object_mask")
        key = tuple((k, tuple(model_input[k].shape)) for k in names) + (tuple(rgb.shape), ) + (
            float(getattr(lf, "alpha", 0.0)), float(getattr(lf, "eikonal_weight", 0.0)), float(getattr(lf, "mask_weight", 0.0)),
            K.get_precision(), K.training_p16(), int(rt.n_steps), K.SCRATCH_GENERATION[0])
        n = model_input["uv"].shape[0] * model_input["uv"].shape[1]
        # host draws, in the order the two-graph path makes them
        u = rt.injected_min_sdf_steps
        if u is None:
            u = torch.empty(int(rt.n_steps)).uniform_(0.0, 1.0, generator=rt.sample_generator)
        eik = model._draw_eikonal(n, dev)
        if self._wgraph is None or key != self._wkey:
            if self._wgraph is not None:
                self._retired_graphs.append((self._wgraph, self._wstatic))
            st_inp = {k: model_input[k].detach().clone() for k in names}
            self._wstatic = (st_inp, eik.clone(), rgb.clone(), torch.empty(int(rt.n_steps), device=dev, dtype=torch.float32))
            st_inp, st_eik, st_rgb, st_u = self._wstatic
            st_u.copy_(u.to(dev).float())
            prev = (rt.use_cuda_graph, rt.injected_min_sdf_steps)
            rt.use_cuda_graph, rt.injected_min_sdf_steps = False, st_u
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):                   # warm-up outside capture (allocator, lazy inits, arena growth)
                        self._trace_and_backward(st_inp, st_eik, st_rgb)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                self._wgraph = torch.cuda.CUDAGraph()
                l0 = K._lib.LAUNCHES[0]
                with torch.cuda.graph(self._wgraph):
                    losses = self._trace_and_backward(st_inp, st_eik, st_rgb)
                    self._wout = {k: v.detach() for k, v in losses.items()}
                self._wlaunches = K._lib.LAUNCHES[0] - l0
            finally:
                rt.use_cuda_graph, rt.injected_min_sdf_steps = prev
            rt._stats["cuda_graph"] = True
            self._wkey = key[:-1] + (K.SCRATCH_GENERATION[0],)
            K.note_graph_captured()
        st_inp, st_eik, st_rgb, st_u = self._wstatic
        for k in names:
            st_inp[k].copy_(model_input[k])
        st_eik.copy_(eik)
        st_rgb.copy_(rgb)
        st_u.copy_(u.to(dev, non_blocking=True).float())
        self._wgraph.replay()
        K._lib.LAUNCHES[0] += self._wlaunches
        return self._wout

    def step(self, model_input, ground_truth) -> torch.Tensor:
        from . import mlp
        b = self.bucket
        model = self.model
        if self._whole_step_supported(model_input):
            losses = self._graphed_step(model_input, ground_truth["rgb"].to(b.flat.device))
        else:
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
        K.sumsq_det(b.grad, self.sumsq, self.sumsq_partials)
        K.clip_adam(b.flat, b.grad, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.t,
                    self.max_norm, self.sumsq, scale)
        mlp.weights_changed()
        self.last_losses = losses
        return losses["loss"]
