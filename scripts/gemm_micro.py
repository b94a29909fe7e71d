"""Micro driver for profiling the contraction kernel: a few representative launches."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idrk import kernels as K

def run(M, N, Kc, prec="3xtf32", mode=K.EPI_SOFTPLUS, reps=3, split_out=True):
    p = K._PRECISION[prec]
    A = torch.randn(M, Kc, device="cuda") * 0.05
    W = torch.randn(N, Kc, device="cuda") * 0.05
    b = torch.randn(N, device="cuda") * 0.01
    A = K.operand(A); W = K.operand(W)
    Ah, Al = K.split_tf32(A) if p == 3 else (A, None)
    Wh, Wl = K.split_tf32(W) if p == 3 else (W, None)
    Ch, Cl = K.empty_padded(M, N, "cuda"), K.empty_padded(M, N, "cuda")
    C = K.empty_padded(M, N, "cuda")
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if split_out and p == 3:
            K.gemm(K.GEMM_NT, Ah, Wh, M, N, Kc, precision=p, A_lo=Al, B_lo=Wl, C_hi=Ch, C_lo=Cl, bias=b, mode=mode, act=100.0)
        else:
            K.gemm(K.GEMM_NT, Ah, Wh, M, N, Kc, precision=p, A_lo=Al, B_lo=Wl, C=C, bias=b, mode=mode, act=100.0)
        e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    print("M=%d N=%d K=%d %s mode=%d split_out=%s: %s us  (%.1f TF/s alg)" % (M, N, Kc, prec, mode, split_out, ["%.1f" % t for t in ts], 2.0 * M * N * Kc / min(ts) / 1e6))

if __name__ == "__main__":
    run(32768, 512, 512)
    run(32768, 512, 32)
    run(32768, 512, 512, mode=K.EPI_NONE, split_out=False)
    run(32768, 512, 512, prec="tf32", split_out=False)
    run(2048, 512, 512)
    run(2048, 512, 512, prec="tf32", split_out=False)


def run_h(M, N, Kc, reps=4, mode=K.EPI_SOFTPLUS, fp32_out=False, dot=False):
    A = torch.randn(M, Kc, device="cuda") * 0.05
    W = torch.randn(N, Kc, device="cuda") * 0.05
    b = torch.randn(N, device="cuda") * 0.01
    Ah, Al = K.split_f16(A); Wh, Wl = K.split_f16(W)
    Ch, Cl = K.empty_half(M, N, "cuda"), K.empty_half(M, N, "cuda")
    C = K.empty_padded(M, N, "cuda")
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if dot:         # fused SDF head: row-dot partials instead of the activation
            K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, bias=b, mode=mode, act=100.0, dot_w=b, dot_out=C)
        elif fp32_out:
            K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, C=C, bias=b, mode=mode, act=100.0)
        else:
            K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, C_h=Ch, C_l=Cl, bias=b, mode=mode, act=100.0)
        e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    print("f16s M=%d N=%d K=%d fp32_out=%s dot=%s: %s us  (%.1f TF/s alg)" % (M, N, Kc, fp32_out, dot, ["%.1f" % t for t in ts], 2.0 * M * N * Kc / min(ts) / 1e6))


if __name__ == "__main__":
    run_h(32768, 512, 512)
    run_h(32768, 512, 64)
    run_h(32768, 512, 512, fp32_out=True)
    run_h(2048, 512, 512)
