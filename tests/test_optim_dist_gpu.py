"""GPU checks of the flat-bucket optimiser (clip + Adam kernel == clip_grad_norm_ + torch.optim.Adam) and of the
DataParallelTrainer step at world size 1."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_clip_adam_matches_torch():
    from idrk import kernels as K
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(100003, generator=g).to(DEV)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    ss = torch.zeros(1, device=DEV)
    for t in range(1, 6):
        grad = (torch.randn(100003, generator=g) * (3.0 if t % 2 else 0.001)).to(DEV)
        pr.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([pr], max_norm=1.0)
        opt.step()
        ss.zero_()
        K.sumsq(grad, ss)
        K.clip_adam(p, grad, m, v, 1e-3, 0.9, 0.999, 1e-8, t, 1.0, ss, 1.0)
        assert torch.allclose(p, pr.detach(), atol=2e-6, rtol=1e-5), t


def test_sumsq_det_is_repeatable_and_correct():
    """The trainer's squared-norm reduction: bit-identical from call to call (replicas must agree on the clip factor),
    equal to a float64 sum within fp32 accumulation error."""
    from idrk import kernels as K
    g = torch.randn(12_345_679, generator=torch.Generator().manual_seed(3)).to(DEV) * 1e-3
    out = torch.zeros(1, device=DEV)
    part = torch.zeros(1184, device=DEV)
    vals = []
    for _ in range(5):
        part.fill_(7.0)
        K.sumsq_det(g, out, part)
        vals.append(out.clone())
    assert all(torch.equal(vals[0], v) for v in vals[1:])
    ref = float((g.double() ** 2).sum())
    assert abs(float(vals[0]) - ref) <= 2e-6 * ref


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainer_step_reduces_loss_and_matches_manual_step(use_graph):
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from oracle import idr_oracle as O
    from tests_support import make_conf, quiet_build
    torch.manual_seed(0)
    conf = make_conf("HashGrid", 6, 5, 64, 512, 1.0, width=128, feature=32)
    model = quiet_build(IDRNetwork, conf).to(DEV).train()
    ref = quiet_build(IDRNetwork, conf).to(DEV).train()
    ref.load_state_dict(model.state_dict())
    inp, rgb = O.synthetic_batch(512, seed=1)
    inp = {k: v.to(DEV) for k, v in inp.items()}
    gt = {"rgb": rgb.to(DEV)}
    eik = torch.rand(256, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1
    u = torch.rand(100, generator=torch.Generator().manual_seed(2))
    for mdl in (model, ref):
        mdl.injected_eikonal_points, mdl.ray_tracer.injected_min_sdf_steps = eik, u
    loss_fn = IDRLoss(0.1, 100.0, 50.0)
    tr = DataParallelTrainer(model, loss_fn, lr=1e-3, max_norm=1.0, world_size=1, use_cuda_graph=use_graph)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    l_a = tr.step(inp, gt)
    lo = loss_fn(ref(inp), gt)
    opt.zero_grad()
    lo["loss"].backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    opt.step()
    assert abs(float(l_a) - float(lo["loss"])) < 1e-5 * max(1.0, abs(float(l_a)))
    for (n1, a), (n2, b) in zip(model.named_parameters(), ref.named_parameters()):
        assert torch.allclose(a, b, atol=5e-6, rtol=1e-4), n1
    first = float(l_a)
    for _ in range(5):
        last = float(tr.step(inp, gt))
    assert last < first


def test_trainer_checkpoint_roundtrip_and_handover_to_torch_adam(tmp_path):
    """Reference-layout checkpoint (utils/checkpoints.py) written by the trainer after 2 steps: (a) a fresh trainer
    restored from it takes the same 3rd step (abs 2e-6); (b) the reference's own optimiser, torch.optim.Adam, loads the
    same optimiser file and takes the same 3rd step (fp32 rounding of the fused kernel: abs 5e-6)."""
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from idrk.utils import checkpoints as ck
    from oracle import idr_oracle as O
    from tests_support import make_conf, quiet_build
    torch.manual_seed(0)
    conf = make_conf("HashGrid", 6, 5, 64, 512, 1.0, width=128, feature=32)
    inp, rgb = O.synthetic_batch(256, seed=1)
    inp = {k: v.to(DEV) for k, v in inp.items()}
    gt = {"rgb": rgb.to(DEV)}
    eik = torch.rand(128, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1
    u = torch.rand(100, generator=torch.Generator().manual_seed(2))
    loss_fn = IDRLoss(0.1, 100.0, 50.0)

    def build():
        m = quiet_build(IDRNetwork, conf).to(DEV).train()
        m.injected_eikonal_points, m.ray_tracer.injected_min_sdf_steps = eik, u
        return m

    model = build()
    tr = DataParallelTrainer(model, loss_fn, lr=1e-3, max_norm=1.0)
    for _ in range(2):
        tr.step(inp, gt)
    root = str(tmp_path / "checkpoints")
    ck.save_checkpoints(root, 2, model, optimizer=tr)

    m2 = build()
    tr2 = DataParallelTrainer(m2, loss_fn, lr=123.0, max_norm=1.0)
    assert ck.load_checkpoints(root, m2, optimizer=tr2) == 2
    assert tr2.t == 2 and tr2.lr == 1e-3 and torch.equal(tr2.m, tr.m) and torch.equal(tr2.v, tr.v)

    m3 = build()
    opt = torch.optim.Adam(m3.parameters(), lr=55.0)
    saved = torch.load(root + "/ModelParameters/latest.pth")
    m3.load_state_dict(saved["model_state_dict"])
    opt.load_state_dict(torch.load(root + "/OptimizerParameters/latest.pth")["optimizer_state_dict"])

    tr.step(inp, gt)
    tr2.step(inp, gt)
    lo = loss_fn(m3(inp), gt)
    opt.zero_grad()
    lo["loss"].backward()
    torch.nn.utils.clip_grad_norm_(m3.parameters(), 1.0)
    opt.step()
    for (n1, a), (_, b), (_, c) in zip(model.named_parameters(), m2.named_parameters(), m3.named_parameters()):
        assert torch.allclose(a, b, atol=2e-6, rtol=1e-5), n1       # not bit-equal: fp32 atomics (split-K, table flush) reorder sums
        assert torch.allclose(a, c, atol=5e-6, rtol=1e-4), n1


def _small_trainer(use_graph, lr=1e-3, alpha=50.0, rays=256):
    from idrk.dist import DataParallelTrainer
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from oracle import idr_oracle as O
    from tests_support import make_conf, quiet_build
    torch.manual_seed(0)
    conf = make_conf("HashGrid", 6, 5, 64, 512, 1.0, width=128, feature=32)
    model = quiet_build(IDRNetwork, conf).to(DEV).train()
    inp, rgb = O.synthetic_batch(rays, seed=1)
    inp = {k: v.to(DEV) for k, v in inp.items()}
    gt = {"rgb": rgb.to(DEV)}
    model.injected_eikonal_points = torch.rand(rays // 2, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1
    model.ray_tracer.injected_min_sdf_steps = torch.rand(100, generator=torch.Generator().manual_seed(2))
    loss_fn = IDRLoss(0.1, 100.0, alpha)
    return DataParallelTrainer(model, loss_fn, lr=lr, max_norm=1.0, world_size=1, use_cuda_graph=use_graph), inp, gt


def _sync(dst, src):
    """Puts trainer `dst` into exactly the state of `src` (parameters, Adam moments, step count): the comparisons below are
    per step from identical states - two free-running fp32 trainers drift apart through atomic-order noise that Adam's
    normalisation amplifies, and a drifting trace eventually flips a ray."""
    from idrk import mlp
    dst.bucket.flat.copy_(src.bucket.flat)
    dst.m.copy_(src.m)
    dst.v.copy_(src.v)
    dst.t = src.t
    mlp.weights_changed()


def _same_step(a, b, inp, gt, step):
    _sync(b, a)
    la, lb = float(a.step(inp, gt)), float(b.step(inp, gt))
    assert abs(la - lb) <= 1e-5 * max(1.0, abs(lb)), (step, la, lb)
    for k in ("rgb_loss", "eikonal_loss", "mask_loss"):
        va, vb = float(a.last_losses[k]), float(b.last_losses[k])
        assert abs(va - vb) <= 1e-5 * max(1e-2, abs(vb)), (step, k, va, vb)
    for (n1, p), (_, q) in zip(a.model.named_parameters(), b.model.named_parameters()):
        assert torch.allclose(p, q, atol=5e-6, rtol=1e-4), (step, n1)


def test_graphed_trainer_follows_alpha_and_lr_changes():
    """The reference loop doubles IDRLoss.alpha at every alpha milestone and decays the lr with MultiStepLR
    (training/idr_train.py:227-228, 131-134): a trainer whose shade + loss + backward is replayed from a CUDA graph must
    pick both up.  A graphed and an eager trainer take 6 steps, each from the same state, with alpha / lr changed after
    step 3: losses rel 1e-5 and updated parameters abs 5e-6 at EVERY step (a stale alpha or lr would miss both by far:
    the mask term scales with alpha, the update with lr)."""
    from idrk.dist import MultiStepLR
    a, inp, gt = _small_trainer(True)
    b, _, _ = _small_trainer(False)
    sa, sb = MultiStepLR(a, [1], gamma=0.5), MultiStepLR(b, [1], gamma=0.5)
    mask_losses = []
    for step in range(6):
        if step == 3:
            a.loss_fn.alpha *= 2.0
            b.loss_fn.alpha *= 2.0
            sa.step()
            sb.step()
            assert a.lr == 0.5e-3 and b.lr == 0.5e-3
        _same_step(a, b, inp, gt, step)
        mask_losses.append(float(a.last_losses["mask_loss"]))
    assert abs(mask_losses[3] - mask_losses[2]) > 1e-4 * abs(mask_losses[2])     # alpha really entered the graphed loss


def test_graphs_survive_growth_of_shared_scratch():
    """SdfPipeline scratch and the zero arena are shared by the tracer's and the trainer's CUDA graphs; a larger SDF
    query after capture (utils.plots.sdf_sweep uses 2^18-row chunks) must neither free memory a graph still points at nor
    leave a stale graph in use: the graphed trainer keeps matching an eager twin step for step."""
    from idrk import kernels as K
    a, inp, gt = _small_trainer(True)
    b, _, _ = _small_trainer(False)
    for step in range(3):                               # capture both graphs
        _same_step(a, b, inp, gt, step)
    gen0 = K.SCRATCH_GENERATION[0]
    pts = torch.rand(200000, 3, device=DEV) * 2 - 1     # far more rows than any tracer query of a 256-ray batch
    with torch.no_grad():
        big = a.model.implicit_network.sdf(pts)
        junk = [torch.randn(1 << 22, device=DEV) for _ in range(8)]      # recycle whatever the allocator got back
    assert K.SCRATCH_GENERATION[0] > gen0
    for step in range(3, 6):
        _same_step(a, b, inp, gt, step)
    del junk, big


def test_trainer_gradients_across_operand_formats_and_leaf_side_path():
    """The flat gradient bucket of one trainer step (eager, same state, same trace inputs) under: the default (bf16 operand
    pairs, weight / bias gradient work on the side stream with per-weight accumulation), the same without the side stream
    (autograd accumulates weight gradients, csrc unchanged), and tf32 hi/lo pairs (round-1 format).  Side stream on / off:
    same arithmetic, other summation order -> 2e-5 of max; bf16 pairs vs tf32 pairs: 1e-3 of max (measured ~2e-5)."""
    from idrk import kernels as K, mlp
    tr, inp, gt = _small_trainer(False, rays=512)
    model = tr.model
    eik = model.injected_eikonal_points.to(DEV)
    traced = model.trace(inp)

    def grads():
        tr._shade_and_backward(traced, eik, gt["rgb"])
        torch.cuda.synchronize()
        return tr.bucket.grad.clone()
    g_side = grads()
    assert g_side.abs().max() > 0
    prev = mlp.LEAF_SIDE.enabled
    try:
        mlp.LEAF_SIDE.enabled = False
        g_main = grads()
    finally:
        mlp.LEAF_SIDE.enabled = prev
    K.set_training_operands("tf32x3")
    try:
        g_tf32 = grads()
    finally:
        K.set_training_operands("p16")
    scale = g_tf32.abs().max()
    assert (g_side - g_main).abs().max() <= 2e-5 * scale
    assert (g_side - g_tf32).abs().max() <= 1e-3 * scale
    # per parameter, relative to that parameter's own largest gradient entry
    for p, o in zip(tr.bucket.params, tr.bucket.offsets):
        a, b = g_side[o:o + p.numel()], g_tf32[o:o + p.numel()]
        if b.abs().max() > 0:
            assert (a - b).abs().max() <= 2e-3 * b.abs().max()


@pytest.mark.parametrize("prec", ["tf32", "fp32"])
def test_trainer_step_in_the_single_pass_precision_modes(prec):
    """set_precision("tf32" | "fp32") (no split operands, hence no side-stream weight-gradient path): a graphed trainer
    step runs, stays finite, and its loss agrees with the default mode's to the mode's accuracy."""
    from idrk import kernels as K
    ref, inp, gt = _small_trainer(True)
    l_ref = float(ref.step(inp, gt))
    K.set_precision(prec)
    try:
        tr, inp2, gt2 = _small_trainer(True)
        l = float(tr.step(inp2, gt2))
        l2 = float(tr.step(inp2, gt2))
        assert torch.isfinite(tr.bucket.flat).all()
    finally:
        K.set_precision("3xtf32")
    assert abs(l - l_ref) <= (5e-2 if prec == "tf32" else 1e-4) * max(1.0, abs(l_ref)), (l, l_ref)
    assert l2 == l2
