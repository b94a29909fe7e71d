import sys, os, torch
sys.path.insert(0, "/root/repo")
import bench
from idrk import kernels as K
from idrk.model.implicit_differentiable_renderer import IDRNetwork
from tests_support import quiet_build
torch.manual_seed(0)
model = quiet_build(IDRNetwork, bench.model_conf()).cuda().eval()
net = model.implicit_network
for n in (1, 100, 128, 129, 4096, 5000, 40000):
    x = (torch.rand(n, 3, device="cuda") * 2 - 1)
    with torch.no_grad():
        K.set_fused_sdf_mlp(False); a = net.sdf(x).clone()
        K.set_fused_sdf_mlp(True); b = net.sdf(x).clone()
    torch.cuda.synchronize()
    print(n, "max diff", (a - b).abs().max().item(), "equal", torch.equal(a, b), flush=True)
# device-count path
x = (torch.rand(4096, 3, device="cuda") * 2 - 1)
cnt = torch.tensor([1000], device="cuda", dtype=torch.int32)
o1 = torch.full((4096,), 9.0, device="cuda"); o2 = o1.clone()
with torch.no_grad():
    K.set_fused_sdf_mlp(False); net.sdf_compacted(x, 4096, cnt, o1)
    K.set_fused_sdf_mlp(True); net.sdf_compacted(x, 4096, cnt, o2)
torch.cuda.synchronize()
print("count path equal", torch.equal(o1[:1000], o2[:1000]), (o2[1024:] == 9.0).all().item())
import time
for npts in (4096, 32700):
  x = (torch.rand(npts, 3, device="cuda") * 2 - 1)
  o1 = torch.empty(npts, device="cuda")
  for flag in (False, True):
    K.set_fused_sdf_mlp(flag)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.no_grad():
        net.sdf_compacted(x, npts, None, o1); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20): net.sdf_compacted(x, npts, None, o1)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): g.replay()
        b.record(); torch.cuda.synchronize()
    print("fused" if flag else "layers", "%.1f us per SDF eval of %d points" % (a.elapsed_time(b) / 100 * 1e3, npts))
