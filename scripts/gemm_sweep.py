"""Device time per launch of the training-path contractions (graph of back-to-back launches, no host gaps)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idrk import kernels as K


def timed(fn, n=40):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g.replay()
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / n * 1e3


def case(layout, M, N, Kc, split_k=1, mode=K.EPI_NONE):
    # operands laid out as the layout expects: NT: A[M,K] B[N,K]; NN: A[M,K] B[K,N]; TN: A[K,M] B[K,N]
    shp_a = (Kc, M) if layout == K.GEMM_TN else (M, Kc)
    shp_b = (N, Kc) if layout == K.GEMM_NT else (Kc, N)
    A = K.operand(torch.randn(*shp_a, device="cuda") * 0.05)
    B = K.operand(torch.randn(*shp_b, device="cuda") * 0.05)
    Ah, Al = K.split_tf32(A); Bh, Bl = K.split_tf32(B)
    C = K.empty_padded(M, N, "cuda")
    b = torch.randn(N, device="cuda")
    def fn():
        K.gemm(layout, Ah, Bh, M, N, Kc, precision=K.PREC_3XTF32, A_lo=Al, B_lo=Bl, C=C, bias=b if mode else None, mode=mode, act=100.0,
               split_k=split_k, accumulate=False)
    us = timed(fn)
    print("%s M=%d N=%d K=%d split=%d: %.1f us  (%.1f TF/s alg)" % ({K.GEMM_NT: "NT", K.GEMM_NN: "NN", K.GEMM_TN: "TN"}[layout], M, N, Kc, split_k, us, 2.0 * M * N * Kc / us / 1e6), flush=True)


if __name__ == "__main__" and len(sys.argv) == 1:
    for M in (2048, 3072, 4096):
        case(K.GEMM_NT, M, 512, 512, mode=K.EPI_SOFTPLUS)
        case(K.GEMM_NN, M, 512, 512)
    for Kc in (2048, 3072):
        for sk in (1, 2, 4, 8):
            case(K.GEMM_TN, 512, 512, Kc, split_k=sk)


def case_h(M, N, Kc):
    A = torch.randn(M, Kc, device="cuda") * 0.05
    W = torch.randn(N, Kc, device="cuda") * 0.05
    b = torch.randn(N, device="cuda") * 0.01
    Ah, Al = K.split_f16(A); Wh, Wl = K.split_f16(W)
    Ch, Cl = K.empty_half(M, N, "cuda"), K.empty_half(M, N, "cuda")
    def fn():
        K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, C_h=Ch, C_l=Cl, bias=b, mode=K.EPI_SOFTPLUS, act=100.0)
    us = timed(fn)
    print("f16s M=%d N=%d K=%d: %.1f us  (%.1f TF/s alg)" % (M, N, Kc, us, 2.0 * M * N * Kc / us / 1e6), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "f16s":
    for M in (2048, 4096, 8192, 32700, 131072):
        case_h(M, 512, 512)
    case_h(32700, 485, 512)
    case_h(32700, 512, 485)
