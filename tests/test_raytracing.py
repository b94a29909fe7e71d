"""GPU parity of the ray tracer.  When both sides see the same SDF values (the analytic SDF is evaluated by
the same host function for the oracle and for the CUDA tracer) masks, distances and points must be BIT-EXACT;
for the golden fixtures, whose SDFs are evaluated with device torch ops that differ from the host in the last
ulp, mismatching rays are counted and bounded."""
import numpy as np
import pytest
import torch

from conftest import assert_flips_borderline
from oracle import idr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
CONF = dict(object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=3,
            sphere_tracing_iters=10, n_steps=100, n_secant_steps=8)


def T(a):
    return torch.from_numpy(np.array(a))


def r2(p):
    return torch.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2])


EXACT_SDFS = {
    "sphere": lambda p: r2(p) - 0.5,
    "bumpy": lambda p: (r2(p) - 0.55) * 0.7 + (p[:, 0] * p[:, 1]) * 0.3 - (p[:, 2] * p[:, 2]) * 0.2,
    "torus": lambda p: torch.sqrt((torch.sqrt(p[:, 0] * p[:, 0] + p[:, 2] * p[:, 2]) - 0.45) *
                                  (torch.sqrt(p[:, 0] * p[:, 0] + p[:, 2] * p[:, 2]) - 0.45) + p[:, 1] * p[:, 1]) - 0.18,
    "slab": lambda p: p[:, 2] * p[:, 2] * 4.0 - 0.05 + p[:, 0] * 0.0,
}


def rays(n, seed, tz=-3.0, f=500.0):
    pose = torch.eye(4).unsqueeze(0)
    pose[0, 2, 3] = tz
    Kc = torch.eye(4).unsqueeze(0)
    Kc[0, 0, 0] = Kc[0, 1, 1] = f
    Kc[0, 0, 2] = Kc[0, 1, 2] = 128.0
    uv = torch.rand(1, n, 2, generator=torch.Generator().manual_seed(seed)) * 256
    mask = torch.rand(1, n, generator=torch.Generator().manual_seed(seed + 1)) > 0.5
    dirs, cam = O.camera_rays(uv, pose, Kc)
    return dirs, cam, mask.reshape(-1)


@pytest.mark.parametrize("kind", list(EXACT_SDFS))
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("n,f", [(777, 500.0), (2048, 150.0)])
def test_bit_exact_vs_oracle(kind, training, n, f):
    from idrk.model.ray_tracing import RayTracing
    dirs, cam, mask = rays(n, 5, f=f)                      # f = 150: many rays miss the unit sphere
    u = torch.rand(100, generator=torch.Generator().manual_seed(9))
    sdf = EXACT_SDFS[kind]
    orc = O.RayTracerOracle(**CONF)
    orc.training = training
    p_ref, m_ref, d_ref = orc(sdf, cam, mask, dirs, u)
    t_sph, hit = O.sphere_intersection(cam, dirs, 1.0)
    tr = RayTracing(**CONF)
    tr.train(training)
    p, m, d = tr(lambda q: sdf(q.cpu()).to(DEV), cam.to(DEV), mask.to(DEV), dirs.to(DEV), min_sdf_steps=u,
                 sphere_intersections=(t_sph.to(DEV), hit.to(DEV)))
    assert torch.equal(m.cpu(), m_ref)
    assert torch.equal(d.cpu(), d_ref)
    assert torch.equal(p.cpu(), p_ref)


@pytest.mark.parametrize("kind", ["sphere", "bumpy", "torus"])
@pytest.mark.parametrize("training", [True, False])
def test_golden_fixtures(golden, kind, training):
    from idrk.model.ray_tracing import RayTracing
    g = golden("raytracing")
    sdfs = {"sphere": lambda p: p.norm(2, dim=1) - 0.5,
            "bumpy": lambda p: (p.norm(2, dim=1) - 0.55) * 0.7 + 0.05 * torch.sin(9.0 * p[:, 0]) * torch.sin(7.0 * p[:, 1]),
            "torus": lambda p: torch.stack([torch.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2) - 0.45, p[:, 1]], 1).norm(2, dim=1) - 0.18}
    tr = RayTracing(**CONF)
    tr.train(training)
    p, m, d = tr(sdfs[kind], T(g["cam"]).to(DEV), T(g["mask"]).reshape(-1).to(DEV), T(g["dirs"]).to(DEV),
                 min_sdf_steps=T(g["min_sdf_steps"]))
    tag = "%s_%s" % (kind, "train" if training else "eval")
    m_ref, d_ref = T(g["net_" + tag]), T(g["dist_" + tag])
    n = m_ref.numel()
    flips = (m.cpu() != m_ref).sum().item()
    assert flips <= max(1, n // 200), flips                 # <= 0.5 % borderline rays
    # ... and only borderline ones: the oracle tracer on the same analytic SDF (CPU libm vs device sin / sqrt: a few ulp
    # of values <= 1) records how close each ray came to deciding differently; flipped rays must be within 2e-6
    orc = O.RayTracerOracle(**CONF)
    orc.training = training
    _, m_orc, _ = orc(sdfs[kind], T(g["cam"]), T(g["mask"]).reshape(-1), T(g["dirs"]), T(g["min_sdf_steps"]))
    assert torch.equal(m_orc, m_ref)                        # the oracle reproduces the reference fixture exactly
    assert_flips_borderline(m, m_ref, orc.margin, 2e-6, tag)
    bad = ((d.cpu() - d_ref).abs() > 1e-4).sum().item()
    assert bad <= max(2, n // 50), bad                      # argmin ties of the 100-sample sweeps


def test_device_count_path_equals_callable_path():
    """The sync-free path (device-side list lengths) and the generic callable path give identical results."""
    from idrk.model.implicit_differentiable_renderer import ImplicitNetwork
    from idrk.model.ray_tracing import RayTracing
    from tests_support import load_sd_into, quiet_build
    cfg = O.EmbedCfg("HashGrid", 6, 5, 2, 64, 512, 1.0)
    sd = O.make_implicit_sd(cfg, torch.Generator().manual_seed(1), perturb=0.02)
    net = quiet_build(ImplicitNetwork, 256, 3, 1, [512] * 8, True, 0.6, (4,), True, 6, "HashGrid", 5, 2, 64, 512, 1.0)
    load_sd_into(net, sd, "implicit_network.")
    net = net.to(DEV)
    dirs, cam, mask = rays(2048, 3)
    u = torch.rand(100, generator=torch.Generator().manual_seed(2))
    tr = RayTracing(**CONF)
    a = tr(net.sdf, cam.to(DEV), mask.to(DEV), dirs.to(DEV), min_sdf_steps=u)
    assert tr.last_stats["fast_path"]
    b = tr(lambda x: net.sdf(x), cam.to(DEV), mask.to(DEV), dirs.to(DEV), min_sdf_steps=u)
    assert not tr.last_stats["fast_path"]
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # and against the oracle tracer driven by the oracle network (fp32 CPU): report-level agreement
    orc = O.RayTracerOracle(**CONF)
    with torch.no_grad():
        p_ref, m_ref, d_ref = orc(lambda x: O.implicit_forward(x, sd, cfg)[:, 0], cam, mask, dirs, u)
    flips = (a[1].cpu() != m_ref).sum().item()
    assert flips <= 2048 // 100, flips
    # MLP SDF through the fp16-pair pipeline vs the fp32 CPU oracle: |sdf difference| <= 1e-5 (asserted below on the
    # oracle's own final points), so only rays within 2e-5 of a decision boundary may flip
    assert_flips_borderline(a[1], m_ref, orc.margin, 2e-5, "MLP sdf")
    with torch.no_grad():
        d_sdf = (net.sdf(p_ref.to(DEV)).cpu() - O.implicit_forward(p_ref, sd, cfg)[:, 0]).abs().max().item()
    assert d_sdf <= 1e-5, d_sdf
    agree = a[1].cpu() == m_ref
    bad = ((a[2].cpu() - d_ref).abs()[agree] > 1e-3).sum().item()
    assert bad <= 2048 // 20, bad


def test_full_size_analytic_sphere_closed_form():
    """BASELINE-sized ray batch (65536 rays): with the exact sphere SDF the traced distance must equal the closed-form
    ray / sphere intersection (size-independent property, no oracle run needed)."""
    from idrk.model.ray_tracing import RayTracing
    n, R = 65536, 0.5
    dirs, cam, mask = rays(n, 11, f=300.0)
    dirs, cam, mask = dirs.to(DEV), cam.to(DEV), mask.to(DEV)
    tr = RayTracing(**CONF)
    tr.eval()
    p, m, d = tr(lambda q: q.norm(2, dim=1) - R, cam, torch.ones_like(mask), dirs)
    dd = dirs.reshape(-1, 3).double()
    c = cam.double()[0]
    b = (dd * c).sum(-1)
    disc = b * b - (c.dot(c) - R * R)
    hit = disc > 1e-6
    t_exact = -b - torch.sqrt(disc.clamp_min(0))
    assert (m[hit]).float().mean().item() > 0.999                 # rays that hit the sphere are classified as hits
    assert (~m[disc < -1e-6]).all()                               # rays that miss are not
    err = (d.double() - t_exact).abs()[hit & m]
    assert err.max().item() < 2e-4, err.max().item()
    assert (p.double() - (c + d.double().unsqueeze(-1) * dd)).abs().max().item() < 1e-5


def test_device_count_path_at_32768_rays():
    """Sync-free CUDA-graph trace == eager device-count trace == callable path on a large batch (MLP SDF)."""
    from idrk.model.implicit_differentiable_renderer import ImplicitNetwork
    from idrk.model.ray_tracing import RayTracing
    from tests_support import load_sd_into, quiet_build
    cfg = O.EmbedCfg("HashGrid", 16, 14, 2, 16, 2048, 1.0)
    sd = O.make_implicit_sd(cfg, torch.Generator().manual_seed(4), perturb=0.02)
    net = quiet_build(ImplicitNetwork, 256, 3, 1, [512] * 8, True, 0.6, (4,), True, 16, "HashGrid", 14, 2, 16, 2048, 1.0)
    load_sd_into(net, sd, "implicit_network.")
    net = net.to(DEV)
    dirs, cam, mask = rays(32768, 8)
    u = torch.rand(100, generator=torch.Generator().manual_seed(6))
    args = (cam.to(DEV), mask.to(DEV), dirs.to(DEV))
    tr = RayTracing(**CONF)
    a = tr(net.sdf, *args, min_sdf_steps=u)
    tr.use_cuda_graph = True
    for _ in range(3):                       # warm-up call, capture call, replay call
        g = tr(net.sdf, *args, min_sdf_steps=u)
    b = tr(lambda x: net.sdf(x), *args, min_sdf_steps=u)
    for x, y, z in zip(a, g, b):
        assert torch.equal(x, y) and torch.equal(x, z)
    assert "n_sampler" in tr.last_stats
