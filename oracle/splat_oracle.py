"""ORACLE (test infrastructure) — 3D Gaussian splatting rasterisation as `gsplat.rendering.rasterization` (gsplat 1.4.0)
performs it for nerfstudio's splatfacto (call site: nerfstudio/models/splatfacto.py:555-581).

PARITY UNPINNED: gsplat's sources are not under /root/reference and the package is not installed, so this file restates the
PUBLISHED algorithm of the pinned version (Kerbl et al. 2023, "3D Gaussian Splatting", eqs. 5-6 and
App. A for the EWA projection) in dense, differentiable torch: every pixel evaluates every Gaussian in depth order, so
`torch.autograd` provides reference gradients for the hand-written backward.  Small cases only (P pixels x N Gaussians).

Stages and conventions (gsplat 1.4, rasterize_mode="classic", packed=False):
  quats (w,x,y,z) are normalised; Sigma = R S S^T R^T; camera point t = V[:3,:3] mu + V[:3,3]; cull t_z outside
  [near, far]; x/z, y/z clamped to 1.3 * tan(fov/2) for the Jacobian; Sigma' = J W Sigma W^T J^T + eps2d I (0.3);
  conic = Sigma'^-1; radius = ceil(3 sqrt(lambda_max)) with lambda_max = b + sqrt(max(0.1, b^2 - det)), b = (a+c)/2;
  mean2d = (fx x/z + cx, fy y/z + cy); colours = max(SH(dir) + 0.5, 0) with dir = normalize(mu - cam_pos), or the given
  [N,3] colours; per pixel centre (x+0.5, y+0.5), front to back: sigma = 0.5 (A dx^2 + C dy^2) + B dx dy (skip if < 0),
  alpha = min(0.999, opacity exp(-sigma)) (skip if < 1/255), stop before a Gaussian that would leave T <= 1e-4.
  A Gaussian only reaches a pixel whose 16x16 tile intersects its radius box.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def quat_to_rotmat(q: Tensor) -> Tensor:
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(*q.shape[:-1], 3, 3)


def sh_basis(degree: int, d: Tensor) -> Tensor:
    """Real SH basis [..., (degree+1)^2] in the 3DGS / gsplat sign convention."""
    x, y, z = d.unbind(-1)
    out = [torch.full_like(x, SH_C0)]
    if degree >= 1:
        out += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if degree >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
        if degree >= 3:
            out += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
                    SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy),
                    SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(out, -1)


def project(means: Tensor, quats: Tensor, scales: Tensor, viewmat: Tensor, K: Tensor, width: int, height: int,
            near: float = 0.01, far: float = 1e10, eps2d: float = 0.3, radius_clip: float = 0.0, tile: int = 16):
    """-> dict(means2d [N,2], depths [N], conics [N,3], radii int [N] (0 = culled), tile_min/max int [N,2])."""
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    pc = means @ R.T + t
    x, y, z = pc.unbind(-1)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    valid = (z > near) & (z < far)
    zs = torch.where(valid, z, torch.ones_like(z))
    lim_x, lim_y = 1.3 * (0.5 * width / fx), 1.3 * (0.5 * height / fy)
    tx = zs * torch.clamp(x / zs, -lim_x, lim_x)
    ty = zs * torch.clamp(y / zs, -lim_y, lim_y)
    zero = torch.zeros_like(zs)
    J = torch.stack([fx / zs, zero, -fx * tx / (zs * zs), zero, fy / zs, -fy * ty / (zs * zs)], -1).reshape(-1, 2, 3)
    Rq = quat_to_rotmat(quats)
    M = Rq * scales[:, None, :]
    cov3 = M @ M.transpose(1, 2)
    T = J @ R
    cov2 = T @ cov3 @ T.transpose(1, 2)
    a, b, c = cov2[:, 0, 0] + eps2d, cov2[:, 0, 1], cov2[:, 1, 1] + eps2d
    det = a * c - b * b
    valid = valid & (det > 0)
    dets = torch.where(valid, det, torch.ones_like(det))
    conics = torch.stack([c / dets, -b / dets, a / dets], -1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radii = torch.ceil(3.0 * torch.sqrt(lam))
    means2d = torch.stack([fx * x / zs + cx, fy * y / zs + cy], -1)
    valid = valid & (radii > radius_clip)
    tiles_x, tiles_y = (width + tile - 1) // tile, (height + tile - 1) // tile
    tmin = torch.stack([torch.clamp(torch.floor((means2d[:, 0] - radii) / tile), 0, tiles_x),
                        torch.clamp(torch.floor((means2d[:, 1] - radii) / tile), 0, tiles_y)], -1)
    tmax = torch.stack([torch.clamp(torch.ceil((means2d[:, 0] + radii) / tile), 0, tiles_x),
                        torch.clamp(torch.ceil((means2d[:, 1] + radii) / tile), 0, tiles_y)], -1)
    valid = valid & ((tmax[:, 0] - tmin[:, 0]) * (tmax[:, 1] - tmin[:, 1]) > 0)
    radii = torch.where(valid, radii, torch.zeros_like(radii)).to(torch.int32)
    return dict(means2d=means2d, depths=z, conics=conics, radii=radii, tile_min=tmin.to(torch.int32), tile_max=tmax.to(torch.int32))


def sh_colors(means: Tensor, viewmat: Tensor, sh: Tensor, degree: int) -> Tensor:
    """colors [N,K,3] (K >= (degree+1)^2) -> view-dependent rgb [N,3] = max(SH + 0.5, 0)."""
    campos = -viewmat[:3, :3].T @ viewmat[:3, 3]
    d = means - campos
    d = d / d.norm(dim=-1, keepdim=True)
    Y = sh_basis(degree, d)
    rgb = (Y[:, :, None] * sh[:, : Y.shape[1], :]).sum(1) + 0.5
    return torch.clamp(rgb, min=0.0)


def rasterization(means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor, viewmat: Tensor,
                  K: Tensor, width: int, height: int, near_plane: float = 0.01, far_plane: float = 1e10,
                  sh_degree: Optional[int] = None, eps2d: float = 0.3, radius_clip: float = 0.0, tile: int = 16,
                  background: Optional[Tensor] = None, with_depth: bool = False) -> Tuple[Tensor, Tensor]:
    """One camera.  -> (render [H,W,3(+1 accumulated depth)], alpha [H,W,1]).  Dense and differentiable."""
    pr = project(means, quats, scales, viewmat, K, width, height, near_plane, far_plane, eps2d, radius_clip, tile)
    rgb = sh_colors(means, viewmat, colors, sh_degree) if sh_degree is not None else colors
    if with_depth:
        rgb = torch.cat([rgb, pr["depths"][:, None]], -1)
    order = torch.argsort(pr["depths"], stable=True)
    order = order[pr["radii"][order] > 0]
    m2, con, op, col = pr["means2d"][order], pr["conics"][order], opacities[order], rgb[order]
    tmin, tmax = pr["tile_min"][order], pr["tile_max"][order]
    ys, xs = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    px, py = xs.reshape(-1).float() + 0.5, ys.reshape(-1).float() + 0.5
    tile_x, tile_y = (xs.reshape(-1) // tile), (ys.reshape(-1) // tile)
    dx = m2[None, :, 0] - px[:, None]
    dy = m2[None, :, 1] - py[:, None]
    sigma = 0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) + con[None, :, 1] * dx * dy
    alpha = torch.clamp(op[None, :] * torch.exp(-sigma), max=0.999)
    in_tile = (tile_x[:, None] >= tmin[None, :, 0]) & (tile_x[:, None] < tmax[None, :, 0]) \
        & (tile_y[:, None] >= tmin[None, :, 1]) & (tile_y[:, None] < tmax[None, :, 1])
    live = in_tile & (sigma >= 0) & (alpha >= 1.0 / 255.0)
    alpha = torch.where(live, alpha, torch.zeros_like(alpha))
    T_after = torch.cumprod(1.0 - alpha, dim=1)                      # transmittance after each Gaussian
    T_before = torch.cat([torch.ones_like(T_after[:, :1]), T_after[:, :-1]], 1)
    stop = (T_after <= 1e-4) & live                                  # the Gaussian that would leave T <= 1e-4 is not blended
    dead = torch.cumsum(stop.to(torch.int32), dim=1) > 0
    w = torch.where(dead, torch.zeros_like(alpha), alpha * T_before)
    out = w @ col
    T_final = 1.0 - w.sum(1)
    if background is not None:
        out = torch.cat([out[:, :3] + T_final[:, None] * background, out[:, 3:]], -1)
    return out.reshape(height, width, -1), (1.0 - T_final).reshape(height, width, 1)
