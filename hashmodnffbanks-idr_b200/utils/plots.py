"""Eval-side consumers of the SDF network with the reference's names (utils/plots.py:110-271): the grids that
marching cubes runs on and the sweep of the SDF over them.

The reference walks a 100^3 ... 512^3 grid in 10 000-point chunks, evaluates the full 257-column network on each
chunk and copies every chunk to the host (`sdf(pnts).detach().cpu().numpy()`, plots.py:116-118, 161-163, 201-203):
one host round trip per chunk.  Here the sweep stays on the device: chunks of 2^18 points go through the SDF-only
inference pipeline (hash / filter-bank encode -> fp16-pair contraction layers -> SDF head; the 256 feature columns
are never formed) and write into ONE device volume that is copied back once.

Mesh extraction itself (skimage.measure.marching_cubes, trimesh export, plotly traces) is host-side I/O and out
of scope; `sdf_volume` returns the array in exactly the layout the reference hands to
`measure.marching_cubes`, together with its `spacing`, so a caller that has skimage continues from there.
"""
import numpy as np
import torch


def _mesh_points(x, y, z, device):
    xx, yy, zz = np.meshgrid(x, y, z)
    pts = torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)
    return pts.to(device) if device is not None else pts


def _default_device():
    return torch.device("cuda") if torch.cuda.is_available() else None


def get_grid_uniform(resolution, device="auto"):
    """[-1, 1]^3 lattice, resolution^3 points in numpy meshgrid('xy') order (plots.py:226-238)."""
    x = np.linspace(-1.0, 1.0, resolution)
    dev = _default_device() if device == "auto" else device
    return {"grid_points": _mesh_points(x, x, x, dev), "shortest_axis_length": 2.0, "xyz": [x, x, x],
            "shortest_axis_index": 0}


def get_grid(points, resolution, device="auto"):
    """Lattice around a point cloud: `resolution` samples along the shortest side of its bounding box (padded by
    eps = 0.2), the same spacing along the other two (plots.py:240-271)."""
    eps = 0.2
    lo = torch.min(points, dim=0)[0].squeeze().cpu().numpy()
    hi = torch.max(points, dim=0)[0].squeeze().cpu().numpy()
    short = int(np.argmin(hi - lo))
    axis = np.linspace(lo[short] - eps, hi[short] + eps, resolution)
    length = np.max(axis) - np.min(axis)
    step = length / (axis.shape[0] - 1)
    xyz = [axis if d == short else np.arange(lo[d] - eps, hi[d] + step + eps, step) for d in range(3)]
    dev = _default_device() if device == "auto" else device
    return {"grid_points": _mesh_points(xyz[0], xyz[1], xyz[2], dev), "shortest_axis_length": length, "xyz": xyz,
            "shortest_axis_index": short}


@torch.no_grad()
def sdf_sweep(sdf, points, chunk=1 << 18, out=None):
    """SDF of every row of `points` [P, 3] -> device tensor [P], evaluated in device-resident chunks.

    `sdf` is an `ImplicitNetwork` (its SDF-only pipeline is used) or any callable [p, 3] -> [p]
    (e.g. the reference's `lambda x: model.implicit_network(x)[:, 0]`)."""
    fn = sdf.sdf if hasattr(sdf, "sdf") and hasattr(sdf, "refresh_inference_weights") else sdf
    P = points.shape[0]
    if out is None:
        out = torch.empty(P, device=points.device, dtype=torch.float32)
    for s in range(0, P, chunk):
        e = min(P, s + chunk)
        out[s:e] = fn(points[s:e]).reshape(-1)
    return out


def sdf_volume(sdf, grid, chunk=1 << 18):
    """(volume, spacing, origin) for `measure.marching_cubes(volume=..., level=0, spacing=...)` as the reference calls
    it (plots.py:122-130): volume[i, j, k] = sdf(x_i, y_j, z_k).  Returns None for the volume when the SDF has no zero
    crossing on the grid (the reference skips meshing then, plots.py:120)."""
    z = sdf_sweep(sdf, grid["grid_points"], chunk)
    x, y, zz = grid["xyz"]
    lo, hi = torch.aminmax(z)
    if lo.item() > 0 or hi.item() < 0:
        return None, None, None
    vol = z.reshape(y.shape[0], x.shape[0], zz.shape[0]).permute(1, 0, 2).contiguous().cpu().numpy().astype(np.float32)
    d = x[2] - x[1]
    return vol, (d, d, d), np.array([x[0], y[0], zz[0]])
