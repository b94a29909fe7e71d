"""Cost of one SDF query of the tracer as a function of the device-side row count (capacity 4096 rows), in a CUDA graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from idrk.model.implicit_differentiable_renderer import IDRNetwork
from tests_support import quiet_build

torch.manual_seed(0)
model = quiet_build(IDRNetwork, bench.model_conf()).cuda().eval()
net = model.implicit_network
x = (torch.rand(4096, 3, device="cuda") * 2 - 1)
out = torch.empty(4096, device="cuda")
for count in (0, 1, 128, 1024, 4096):
    cnt = torch.tensor([count], device="cuda", dtype=torch.int32)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.no_grad():
        net.sdf_compacted(x, 4096, cnt, out); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                net.sdf_compacted(x, 4096, cnt, out)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g.replay()
        b.record(); torch.cuda.synchronize()
    print("count %5d: %.1f us per query" % (count, a.elapsed_time(b) / 100 * 1e3), flush=True)
