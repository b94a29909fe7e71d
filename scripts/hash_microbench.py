"""Hash-encode microbench (BASELINE cfg5): fwd / bwd Mpts/s and fraction of the HBM roofline."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def time_cuda(fn, warm=3, iters=10, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def run(n, log2T, mode, L=16, F=2, flush=None):
    from idrk import kernels as K
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    m = MultiResHashGridMLP(True, 3, L, F, log2T, 16, 2048, frac_mode=mode).cuda()
    spec, tables, B = m.spec(), tuple(t.detach() for t in m.tables()), m.freq_encoding.B
    x = torch.rand(n, 3, device="cuda")
    out = torch.empty(n, K.pad4(spec.width), device="cuda")
    dy = torch.randn(n, K.pad4(spec.width), device="cuda")
    grads = [torch.zeros_like(t) for t in tables]
    G = 1 if mode == "reference" else 8
    t_f = time_cuda(lambda: K.hash_encode_fwd(spec, x, tables, B, out=out), flush=flush)
    t_b = time_cuda(lambda: K.hash_encode_bwd(spec, x, tables, B, dy, grads, False), flush=flush)      # table gradients
    t_bx = time_cuda(lambda: K.hash_encode_bwd(spec, x, tables, B, dy, grads, True), flush=flush)     # + dL/dx
    # algorithmic bytes per point.  fwd: x + prefix columns + level columns written, one F-float table row read per
    # gather.  bwd (table gradients): x + the level columns of dL/dy read (the prefix columns are not needed), one
    # read-modify-write of a table row per reduction.  bwd + dL/dx additionally reads the prefix columns of dL/dy and
    # the table rows, and writes dx.
    pre = 4 * (3 + 2 * spec.n_fourier)
    bf = 12 + pre + 4 * L * F + G * L * 4 * F
    bb = 12 + 4 * L * F + 2 * G * L * 4 * F
    bbx = bb + pre + 12 + (G * L * 4 * F if mode != "reference" else 0)
    hbm, src = peaks()
    return {"n": n, "log2T": log2T, "mode": mode,
            "fwd_ms": t_f, "fwd_mpts": n / t_f / 1e3, "fwd_frac": n * bf / (t_f * 1e-3) / (hbm * 1e9),
            "bwd_ms": t_b, "bwd_mpts": n / t_b / 1e3, "bwd_frac": n * bb / (t_b * 1e-3) / (hbm * 1e9),
            "bwd_with_dx_ms": t_bx, "bwd_with_dx_mpts": n / t_bx / 1e3,
            "bwd_with_dx_frac": n * bbx / (t_bx * 1e-3) / (hbm * 1e9),
            "bytes_per_pt": [bf, bb, bbx], "peak": src}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[1 << 20, 1 << 24])
    ap.add_argument("--log2T", type=int, nargs="+", default=[14, 19, 22])
    ap.add_argument("--modes", nargs="+", default=["reference", "trilinear"])
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mode in a.modes:
        for lt in a.log2T:
            for n in a.n:
                print(json.dumps(run(n, lt, mode, flush=flush)), flush=True)
