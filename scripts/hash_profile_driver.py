"""ncu driver: one launch of every hash-encode kernel variant (4M points, L=16, F=2, T=2^19 or argv[1]).

    python scripts/hash_profile_driver.py 19 > gpurun_out/plain_hash.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:"hash_encode|rs_|morton" -o gpurun_out/hash_all \
        python scripts/hash_profile_driver.py 19

Launch order per mode (reference, then trilinear): forward, table-gradient backward, backward + dL/dx; trilinear then
adds the Z-order walk of the SAME unordered batch: the Morton radix sort (key kernel + 3 x [hist, scan, scan, scatter]),
forward through the permutation, table-gradient backward through the permutation (run-aggregating instantiation).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idrk import kernels as K                                                   # noqa: E402
from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP          # noqa: E402

log2T = int(sys.argv[1]) if len(sys.argv) > 1 else 19
n = 1 << (int(os.environ.get("IDRK_PROFILE_LOG2N", "22")))          # 2^22 points for `--set full` (replays), 2^24 for a light metric pass
modes = sys.argv[2:] if len(sys.argv) > 2 else ["reference", "trilinear"]
for mode in modes:
    m = MultiResHashGridMLP(True, 3, 16, 2, log2T, 16, 2048, frac_mode=mode).cuda()
    spec, tables, B = m.spec(), tuple(t.detach() for t in m.tables()), m.freq_encoding.B
    x = torch.rand(n, 3, device="cuda")
    out = torch.empty(n, K.pad4(spec.width), device="cuda")
    dy = torch.randn(n, K.pad4(spec.width), device="cuda")
    grads = [torch.zeros_like(t) for t in tables]
    K.hash_encode_fwd(spec, x, tables, B, out=out)
    K.hash_encode_bwd(spec, x, tables, B, dy, grads, False)
    K.hash_encode_bwd(spec, x, tables, B, dy, grads, True)
    if mode != "reference":
        perm = K.morton_perm(x)
        K.hash_encode_fwd(spec, x, tables, B, out=out, perm=perm)
        K.hash_encode_bwd(spec, x, tables, B, dy, grads, False, perm=perm)
    torch.cuda.synchronize()
    print(mode, "ok", float(out[0, 40]))
