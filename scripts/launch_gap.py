"""Per-node cost of a chain of dependent tiny kernels inside a CUDA graph (launch-gap floor of the step graphs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idrk import kernels as K

x = torch.randn(1024, 8, device="cuda")
hi, lo = torch.empty_like(x), torch.empty_like(x)
def chain(n):
    for _ in range(n):
        K.split_into(x, 1024, 8, 1.0, hi, lo, 8, 0, None)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain(10)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        chain(1000)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print("graph: %.2f us per tiny dependent kernel" % (a.elapsed_time(b) / 10 / 1000 * 1e3))
