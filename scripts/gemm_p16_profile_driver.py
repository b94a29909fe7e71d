"""Driver for an ncu capture of the 16-bit-pair training contraction (csrc/gemm_p16.cu) at the shapes of one cfg2 step:
    ncu --set full --clock-control none --import-source on -k regex:gemm_p16 -o gpurun_out/gemm_p16 python scripts/gemm_p16_profile_driver.py
NT forward layer with softplus + derivative + pair output (M = 4096), NN input gradient, TN weight gradient (split-K 4).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idrk import kernels as K          # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
M, N, Kc = 4096, 512, 512
X = torch.randn(M, Kc, device=dev, generator=g) * 0.1
W = torch.randn(N, Kc, device=dev, generator=g) * 0.05
dZ = torch.randn(M, N, device=dev, generator=g) * 1e-4
b = torch.zeros(N, device=dev)
xp, wp, dp = K.split_p16(X, 1), K.split_p16(W, 1), K.split_p16(dZ, 1)
for _ in range(2):
    H, S = K.empty_padded(M, N, dev), K.empty_padded(M, N, dev)
    Ch, Cl = K.empty_pair16(M, N, dev, 1)
    K.gemm_p16(K.GEMM_NT, xp, wp, M, N, Kc, C=H, C_pair=(Ch, Cl, 1), S=S, bias=b, mode=K.EPI_SOFTPLUS, act=100.0)   # forward layer
    dX = K.empty_padded(M, Kc, dev)
    K.gemm_p16(K.GEMM_NN, dp, wp, M, Kc, N, C=dX)                                                                    # dX = dZ W
    dW = torch.zeros(N, Kc, device=dev)
    K.gemm_p16(K.GEMM_TN, dp, xp, N, Kc, M, C=dW, split_k=4)                                                         # dW = dZ^T X
torch.cuda.synchronize()
