"""ncu driver: the fp16-pair contraction at the bench's two shapes (100-sample sweep chunk, sphere-tracing batch).

    ncu --set full --clock-control none --import-source on -k regex:gemm_f16s -o gpurun_out/gemm_f16s python scripts/gemm_f16s_profile_driver.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.gemm_micro import run_h          # noqa: E402

run_h(32700, 512, 512, reps=2)
run_h(4096, 512, 512, reps=2)
run_h(32700, 512, 512, reps=2, dot=True)        # last hidden layer with the fused SDF-head row-dot
run_h(32700, 512, 512, reps=2, fp32_out=True)   # what that layer did before (fp32 activation stored, head reads it back)
torch.cuda.synchronize()
