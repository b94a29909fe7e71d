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


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def time_cuda(fn, warm=3, iters=10, flush=None):
    """Median device time of fn (CUDA events on the launching stream).  With torch.distributed initialised every
    iteration starts behind a barrier and counts as the MAX over ranks."""
    import torch.distributed as dist
    world = _world()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        if world > 1:
            dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    if world > 1:
        t = torch.tensor(ts, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts = t.tolist()
    ts.sort()
    return ts[len(ts) // 2]


CHUNK = 1 << 26      # points per launch: N = 256 M is processed in place, 3 GB of input, output rows recycled per chunk


def run(n, log2T, mode, L=16, F=2, flush=None):
    from idrk import kernels as K
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    m = MultiResHashGridMLP(True, 3, L, F, log2T, 16, 2048, frac_mode=mode).cuda()
    spec, tables, B = m.spec(), tuple(t.detach() for t in m.tables()), m.freq_encoding.B
    x_all = torch.rand(n, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    nc = min(n, CHUNK)
    xs = list(x_all.split(nc))
    x = xs[0]
    out = torch.empty(nc, K.pad4(spec.width), device="cuda")
    dy = torch.randn(nc, K.pad4(spec.width), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    # gradient tables = views of ONE flat buffer (what the trainer's bucket looks like, and what NCCL reduces)
    flat = torch.zeros(sum(t.numel() for t in tables), device="cuda")
    grads, o = [], 0
    for t in tables:
        grads.append(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
    G = 1 if mode == "reference" else 8
    world = _world()
    if world > 1:
        import torch.distributed as dist

        def bwd(want_dx):
            # cfg5 at N > 1: every rank owns n points, the step ends with the table-gradient all-reduce
            for xc in xs:
                K.hash_encode_bwd(spec, xc, tables, B, dy[:xc.shape[0]], grads, want_dx)
            dist.all_reduce(flat)
    else:
        def bwd(want_dx, ordered=False):
            for xc in xs:
                K.hash_encode_bwd(spec, xc, tables, B, dy[:xc.shape[0]], grads, want_dx, ordered=ordered)

    def fwd():
        for xc in xs:
            K.hash_encode_fwd(spec, xc, tables, B, out=out[:xc.shape[0]])
    t_f = time_cuda(fwd, flush=flush)
    t_b = time_cuda(lambda: bwd(False), flush=flush)      # table gradients
    t_bx = time_cuda(lambda: bwd(True), flush=flush)      # + dL/dx
    extra = {}
    if world == 1 and len(xs) == 1:
        # Spatially ordered walks.  (a) "presorted": the batch already IS in Z-order (what ray-marched samples look like
        # to the encoder) - nothing is timed but the passes.  (b) "perm": the batch stays as given (uniform random) and the
        # passes walk it through a permutation from the library's own Morton radix sort; reported without the sort, with
        # the sort inside each pass's time, and as a forward + backward pair that shares one sort (a training step).
        bounds = dict(lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0))
        perm = K.morton_perm(x, **bounds)
        x_sorted = x[perm.long()].contiguous()
        xs_saved, xs[0] = xs[0], x_sorted
        t_bs = time_cuda(lambda: bwd(False, ordered=True), flush=flush)
        t_fs = time_cuda(fwd, flush=flush)
        xs[0] = xs_saved
        del x_sorted
        t_sort = time_cuda(lambda: K.morton_perm(x, **bounds), flush=flush)
        t_fp = time_cuda(lambda: K.hash_encode_fwd(spec, x, tables, B, out=out, perm=perm), flush=flush)
        t_bp = time_cuda(lambda: K.hash_encode_bwd(spec, x, tables, B, dy, grads, False, perm=perm), flush=flush)

        def pair():
            pm = K.morton_perm(x, **bounds)
            K.hash_encode_fwd(spec, x, tables, B, out=out, perm=pm)
            K.hash_encode_bwd(spec, x, tables, B, dy, grads, False, perm=pm)

        def pair_plain():
            K.hash_encode_fwd(spec, x, tables, B, out=out)
            K.hash_encode_bwd(spec, x, tables, B, dy, grads, False)
        t_pair = time_cuda(pair, flush=flush)
        t_pair_plain = time_cuda(pair_plain, flush=flush)
        extra = {"bwd_presorted_ms": t_bs, "fwd_presorted_ms": t_fs, "sort_ms": t_sort, "fwd_perm_ms": t_fp, "bwd_perm_ms": t_bp,
                 "pair_sorted_ms": t_pair, "pair_plain_ms": t_pair_plain}
        del perm
    n = n * world                                         # whole-job points per pass (weak scaling: n per GPU)
    # algorithmic bytes per point.  fwd: x + prefix columns + level columns written, one F-float table row read per
    # gather.  bwd (table gradients): x + the level columns of dL/dy read (the prefix columns are not needed), one
    # read-modify-write of a table row per reduction.  bwd + dL/dx additionally reads the prefix columns of dL/dy and
    # the table rows, and writes dx.
    pre = 4 * (3 + 2 * spec.n_fourier)
    bf = 12 + pre + 4 * L * F + G * L * 4 * F
    bb = 12 + 4 * L * F + 2 * G * L * 4 * F
    bbx = bb + pre + 12 + (G * L * 4 * F if mode != "reference" else 0)
    hbm, src = peaks()
    if extra:
        e = extra

        def rate(ms, nbytes):
            return n / ms / 1e3, n * nbytes / (ms * 1e-3) / (hbm * 1e9)
        for key, ms, nb in (("bwd_presorted", e["bwd_presorted_ms"], bb), ("fwd_presorted", e["fwd_presorted_ms"], bf),
                            ("fwd_perm", e["fwd_perm_ms"], bf), ("bwd_perm", e["bwd_perm_ms"], bb),
                            ("fwd_sort_included", e["fwd_perm_ms"] + e["sort_ms"], bf),
                            ("bwd_sort_included", e["bwd_perm_ms"] + e["sort_ms"], bb),
                            ("pair_one_sort", e["pair_sorted_ms"], bf + bb), ("pair_plain", e["pair_plain_ms"], bf + bb)):
            e[key + "_mpts"], e[key + "_frac"] = rate(ms, nb)
    return {**extra, "n": n, "log2T": log2T, "mode": mode, "n_gpus": world, "hbm_peak_gbs_all_gpus": hbm * world,
            "fwd_ms": t_f, "fwd_mpts": n / t_f / 1e3, "fwd_frac": n * bf / (t_f * 1e-3) / (hbm * world * 1e9),
            "bwd_ms": t_b, "bwd_mpts": n / t_b / 1e3, "bwd_frac": n * bb / (t_b * 1e-3) / (hbm * world * 1e9),
            "bwd_with_dx_ms": t_bx, "bwd_with_dx_mpts": n / t_bx / 1e3,
            "bwd_with_dx_frac": n * bbx / (t_bx * 1e-3) / (hbm * world * 1e9),
            "bytes_per_pt": [bf, bb, bbx], "peak": src}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[1 << 20, 1 << 24])
    ap.add_argument("--log2T", type=int, nargs="+", default=[14, 19, 22])
    ap.add_argument("--modes", nargs="+", default=["reference", "trilinear"])
    a = ap.parse_args()
    rank = 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:          # torchrun: cfg5 at 2 / 4 / 8 GPUs, --n is per GPU
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
        rank = dist.get_rank()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mode in a.modes:
        for lt in a.log2T:
            for n in a.n:
                r = run(n, lt, mode, flush=flush)
                if rank == 0:
                    print(json.dumps(r), flush=True)
    if _world() > 1:
        torch.distributed.destroy_process_group()
