"""Device-side row counts of the SDF queries of one trace (bench workload)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from idrk.model.implicit_differentiable_renderer import IDRNetwork
from tests_support import synthetic_batch
from tests_support import quiet_build

torch.manual_seed(0)
model = quiet_build(IDRNetwork, bench.model_conf()).cuda().train()
inp, rgb = synthetic_batch(bench.N_RAYS, seed=1)
inp = {k: v.cuda() for k, v in inp.items()}
for _ in range(2):
    model.trace(inp)
torch.cuda.synchronize()
T = list(model.ray_tracer._states.values())[0]
print("counters:", T.counters.cpu().tolist())
