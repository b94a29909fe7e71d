"""Camera / ray helpers with the reference's names (utils/rend_util.py:48-162).  Image loaders and
COLMAP helpers of the reference file are I/O and out of scope.  These are O(rays) elementwise device
ops that stay differentiable w.r.t. the pose (needed by --train_cameras)."""
import torch
from torch.nn import functional as F


def quat_to_rot(q):
    q = F.normalize(q, dim=1)
    qr, qi, qj, qk = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [1 - 2 * (qj ** 2 + qk ** 2), 2 * (qj * qi - qk * qr), 2 * (qi * qk + qr * qj),
            2 * (qj * qi + qk * qr), 1 - 2 * (qi ** 2 + qk ** 2), 2 * (qj * qk - qi * qr),
            2 * (qk * qi - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * (qi ** 2 + qj ** 2)]
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


def lift(x, y, z, intrinsics):
    K = intrinsics.to(x.device)
    fx, fy = K[:, 0, 0, None], K[:, 1, 1, None]
    cx, cy, sk = K[:, 0, 2, None], K[:, 1, 2, None], K[:, 0, 1, None]
    x_lift = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    y_lift = (y - cy) / fy * z
    return torch.stack((x_lift, y_lift, z, torch.ones_like(z)), dim=-1)


def _pose44(pose):
    if pose.shape[1] != 7:
        return pose
    p = torch.eye(4, device=pose.device, dtype=pose.dtype).repeat(pose.shape[0], 1, 1)
    p[:, :3, :3] = quat_to_rot(pose[:, :4])
    p[:, :3, 3] = pose[:, 4:]
    return p


def _kernel_path(uv, pose, intrinsics):
    """The one-launch kernel (csrc/render_glue.cu) serves fixed cameras on the device; trainable poses (--train_cameras
    needs d ray / d pose) and anything else keep the differentiable tensor-op formulation below."""
    return (uv.is_cuda and uv.dtype == torch.float32 and pose.dtype == torch.float32 and intrinsics.dtype == torch.float32
            and not (torch.is_grad_enabled() and (pose.requires_grad or uv.requires_grad or intrinsics.requires_grad)))


def camera_rays_and_sphere(uv, pose, intrinsics, r=1.0):
    """get_camera_params + get_sphere_intersection in one launch: (ray_dirs, cam_loc, t_sph [B,N,2], hit [B,N])."""
    from .. import kernels as K
    return K.camera_rays(uv, _pose44(pose), intrinsics.to(uv.device), radius=r)


def get_camera_params(uv, pose, intrinsics):
    """uv [B,N,2], pose [B,4,4] (or [B,7] quaternion + location) -> ray_dirs [B,N,3], cam_loc [B,3]."""
    if _kernel_path(uv, pose, intrinsics):
        from .. import kernels as K
        return K.camera_rays(uv, _pose44(pose), intrinsics.to(uv.device))
    if pose.shape[1] == 7:
        cam_loc = pose[:, 4:]
        p = torch.eye(4, device=pose.device, dtype=pose.dtype).repeat(pose.shape[0], 1, 1)
        p[:, :3, :3] = quat_to_rot(pose[:, :4])
        p[:, :3, 3] = cam_loc
    else:
        cam_loc = pose[:, :3, 3]
        p = pose
    b = uv.shape[0]
    x_cam = uv[:, :, 0].view(b, -1)
    y_cam = uv[:, :, 1].view(b, -1)
    z_cam = torch.ones_like(x_cam)
    pix = lift(x_cam, y_cam, z_cam, intrinsics=intrinsics).permute(0, 2, 1)
    world = torch.bmm(p, pix).permute(0, 2, 1)[:, :, :3]
    ray_dirs = F.normalize(world - cam_loc[:, None, :], dim=2)
    return ray_dirs, cam_loc


def get_sphere_intersection(cam_loc, ray_directions, r=1.0):
    """Near/far ray parameters on the bounding sphere, [B,N,2], and the hit mask [B,N]."""
    n_imgs, n_pix, _ = ray_directions.shape
    dot = torch.bmm(ray_directions, cam_loc.unsqueeze(-1)).squeeze(-1)
    under = (dot ** 2 - (cam_loc.norm(2, 1, keepdim=True) ** 2 - r ** 2)).reshape(-1)
    hit = under > 0
    t = torch.zeros(n_imgs * n_pix, 2, device=ray_directions.device)
    root = torch.sqrt(under[hit]).unsqueeze(-1)
    t[hit] = root * torch.tensor([-1.0, 1.0], device=t.device) - dot.reshape(-1)[hit].unsqueeze(-1)
    return t.reshape(n_imgs, n_pix, 2).clamp_min(0.0), hit.reshape(n_imgs, n_pix)


def get_depth(points, pose):
    b, n, _ = points.shape
    if pose.shape[1] == 7:
        full = torch.eye(4, device=points.device).unsqueeze(0).repeat(b, 1, 1)
        full[:, :3, 3] = pose[:, 4:]
        full[:, :3, :3] = quat_to_rot(pose[:, :4])
        pose = full
    hom = torch.cat((points, torch.ones((b, n, 1), device=points.device)), dim=2).permute(0, 2, 1)
    return torch.inverse(pose).bmm(hom)[:, 2, :][:, :, None]
