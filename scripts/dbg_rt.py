import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import idr_oracle as O
from test_raytracing import EXACT_SDFS, rays, CONF
from idrk.model.ray_tracing import RayTracing
for kind, training, n, f in [("sphere", True, 777, 500.0), ("torus", False, 777, 500.0), ("bumpy", False, 2048, 150.0)]:
    dirs, cam, mask = rays(n, 5, f=f)
    u = torch.rand(100, generator=torch.Generator().manual_seed(9))
    sdf = EXACT_SDFS[kind]
    orc = O.RayTracerOracle(**CONF); orc.training = training
    p_ref, m_ref, d_ref = orc(sdf, cam, mask, dirs, u)
    t_sph, hit = O.sphere_intersection(cam, dirs, 1.0)
    tr = RayTracing(**CONF); tr.train(training)
    p, m, d = tr(sdf, cam.cuda(), mask.cuda(), dirs.cuda(), min_sdf_steps=u, sphere_intersections=(t_sph.cuda(), hit.cuda()))
    dd = (d.cpu() - d_ref).abs()
    bad = torch.nonzero(dd > 0).flatten()
    print(kind, training, n, "stats", orc.stats, tr.last_stats, "nbad", bad.numel(), "maxdiff", dd.max().item())
    for i in bad[:6].tolist():
        print("  ray", i, "d", d[i].item(), "ref", d_ref[i].item(), "m", bool(m[i]), bool(m_ref[i]), "obj", bool(mask[i]), "hit", bool(hit.reshape(-1)[i]))
    # check sdf equality host vs device on random pts
    q = torch.rand(10000, 3) * 2 - 1
    print("  sdf host==dev:", torch.equal(sdf(q), sdf(q.cuda()).cpu()))
