"""Spatial ordering of point batches for the 8-corner hash-grid backward.

The table-gradient scatter (csrc/hash_encode.cu, K2) merges consecutive points of a lane that fall into the same cell of
the lane's level before issuing reductions.  Ray-marched samples arrive spatially ordered by construction; for unordered
batches `morton_order` returns the permutation that sorts the points along a Z-order curve (keys are formed with a few
elementwise device ops, the key sort is torch.sort - a library radix sort, host-side plumbing and not part of the
kernel path)."""
import torch


def _spread_bits_10(v: torch.Tensor) -> torch.Tensor:
    """Inserts two zero bits between each of the low 10 bits of an int32 tensor."""
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


@torch.no_grad()
def morton_keys(x: torch.Tensor, lo=None, hi=None) -> torch.Tensor:
    """30-bit Z-order keys of [n, 3] points on a 1024^3 lattice over their bounding box (or [lo, hi])."""
    x = x[:, :3]
    lo = x.min(0).values if lo is None else torch.as_tensor(lo, device=x.device, dtype=x.dtype)
    hi = x.max(0).values if hi is None else torch.as_tensor(hi, device=x.device, dtype=x.dtype)
    q = ((x - lo) / (hi - lo).clamp_min(1e-30) * 1023.0).clamp_(0.0, 1023.0).int()
    return _spread_bits_10(q[:, 0]) | (_spread_bits_10(q[:, 1]) << 1) | (_spread_bits_10(q[:, 2]) << 2)


@torch.no_grad()
def morton_order(x: torch.Tensor, lo=None, hi=None) -> torch.Tensor:
    """Permutation (int64 [n]) that puts the points in Z-order: x[perm], dy[perm] feed the backward."""
    return torch.sort(morton_keys(x, lo, hi)).indices
