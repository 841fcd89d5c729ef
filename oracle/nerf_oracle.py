"""CPU oracle for the volumetric-rendering hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, function-style restatement (torch, CPU, fp32) of the
algorithms the reference (nerfstudio v1.1.5, `/root/reference`) runs on its pure-PyTorch
path for the rows of SURVEY.md §8(a).  It is never imported by the product package
`nerfstudio_b200`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` leg may use it, and only as the checker / the timed
CPU baseline.

Pinning: `tests/golden/make_golden.py` (run in the build container, where the reference
is importable) records inputs and outputs of the *actual* reference modules into
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against
those vectors.  The nerfacc / tiny-cuda-nn semantics (packed path, tcnn-mode grid) are
NOT available as source or wheels: those functions restate the published behaviour of
nerfacc 0.5.2 / tcnn @b3473c8 as recalled from their call sites and are marked
"parity unpinned" individually.

Every function cites the reference file:line it follows (paths relative to
`/root/reference/nerfstudio/`).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

HASH_PRIMES = (1, 2654435761, 805459861)  # field_components/encodings.py:410


# ----------------------------------------------------------------------------------------
# hash grid (a11)
# ----------------------------------------------------------------------------------------
def hash_level_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """Per-level scale, exactly as the reference builds it (encodings.py:339-342).

    The pow is evaluated by torch on a LongTensor exponent with a python-float base, i.e.
    in float32 — which is why the last nerfacto level is 2047, not 2048.
    """
    levels = torch.arange(num_levels)
    growth = math.exp((math.log(max_res) - math.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    import numpy as np  # the reference's base is an np.float64; keep the same dispatch

    return torch.floor(min_res * np.float64(growth) ** levels)


def hash_corner_indices(x: Tensor, scalings: Tensor, log2_T: int) -> Tuple[Tensor, Tensor]:
    """Indices of the 8 corners per (point, level) and the lerp offsets.

    Follows encodings.py:398-435.  Corner order k=0..7 is the reference's `hashed_0..7`:
      0:(c,c,c) 1:(c,f,c) 2:(f,f,c) 3:(f,c,c) 4:(c,c,f) 5:(c,f,f) 6:(f,f,f) 7:(f,c,f)
    Returns (idx int64 [N,L,8] including the per-level offset l*T, offset fp32 [N,L,3]).
    """
    T = 1 << log2_T
    L = scalings.numel()
    scaled = x[:, None, :].float() * scalings.view(-1, 1).float()  # [N,L,3]
    c = torch.ceil(scaled).to(torch.int32).to(torch.int64)
    f = torch.floor(scaled).to(torch.int32).to(torch.int64)
    offset = scaled - torch.floor(scaled).to(torch.int32)
    pick = ((c, c, c), (c, f, c), (f, f, c), (f, c, c), (c, c, f), (c, f, f), (f, f, f), (f, c, f))
    level_off = (torch.arange(L, dtype=torch.int64) * T).view(1, L)
    out = []
    for px, py, pz in pick:
        h = (px[..., 0] * HASH_PRIMES[0]) ^ (py[..., 1] * HASH_PRIMES[1]) ^ (pz[..., 2] * HASH_PRIMES[2])
        out.append(h % T + level_off)
    return torch.stack(out, dim=-1), offset


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    """Multiresolution hash encoding, torch-mode (encodings.py:417-458).  [N,3]->[N,L*F]."""
    idx, o = hash_corner_indices(x, scalings, log2_T)
    f = [table[idx[..., k]] for k in range(8)]  # each [N,L,F]
    ox, oy, oz = o[..., 0:1], o[..., 1:2], o[..., 2:3]
    f03 = f[0] * ox + f[3] * (1 - ox)
    f12 = f[1] * ox + f[2] * (1 - ox)
    f56 = f[5] * ox + f[6] * (1 - ox)
    f47 = f[4] * ox + f[7] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    enc = f0312 * oz + f4756 * (1 - oz)
    return enc.flatten(-2, -1)


# --- tcnn-mode grid ("parity unpinned": tiny-cuda-nn @ b3473c8 is not available) --------
def tcnn_grid_meta(num_levels: int, base_res: int, per_level_scale: float, log2_T: int, F: int = 2):
    """Level table of tcnn's HashGrid as published: scale, res, offset, hashed."""
    meta, offset = [], 0
    for l in range(num_levels):
        scale = math.exp2(l * math.log2(per_level_scale)) * base_res - 1.0
        res = int(math.ceil(scale)) + 1
        n = res ** 3
        n = (n + 7) // 8 * 8
        size = min(n, 1 << log2_T)
        meta.append(dict(scale=float(scale), res=res, offset=offset, size=size, hashed=res ** 3 > size))
        offset += size
    return meta, offset


def tcnn_hash_encode(x: Tensor, table: Tensor, meta) -> Tensor:
    """tcnn-mode forward: pos = x*scale+0.5, dense index on coarse levels, uint32 hash else."""
    outs = []
    for m in meta:
        pos = x.float() * m["scale"] + 0.5
        g = torch.floor(pos)
        w = pos - g
        g = g.to(torch.int64)
        acc = 0
        for corner in range(8):
            d = [(corner >> a) & 1 for a in range(3)]
            cg = torch.stack([g[:, a] + d[a] for a in range(3)], -1)
            wt = 1.0
            for a in range(3):
                wt = wt * (w[:, a] if d[a] else 1 - w[:, a])
            if m["hashed"]:
                h = ((cg[:, 0] * HASH_PRIMES[0]) & 0xFFFFFFFF) ^ ((cg[:, 1] * HASH_PRIMES[1]) & 0xFFFFFFFF) ^ (
                    (cg[:, 2] * HASH_PRIMES[2]) & 0xFFFFFFFF
                )
                idx = h % m["size"]
            else:
                r = m["res"]
                idx = (cg[:, 0] + cg[:, 1] * r + cg[:, 2] * r * r) % m["size"]
            acc = acc + wt[:, None] * table[idx + m["offset"]].float()
        outs.append(acc)
    return torch.cat(outs, -1)


# ----------------------------------------------------------------------------------------
# direction / frequency encodings (a17, a28)
# ----------------------------------------------------------------------------------------
def sh_components(levels: int, d: Tensor) -> Tensor:
    """Real SH basis, positive-sign convention (utils/spherical_harmonics.py:24-81).

    The torch path of SHEncoding evaluates this directly on the [0,1]-mapped directions
    (encodings.py:791-794; fields/base_field.py:136-142).
    """
    deg = levels - 1
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x ** 2, y ** 2, z ** 2
    c = [torch.full_like(x, 0.28209479177387814)]
    if deg > 0:
        c += [0.4886025119029199 * y, 0.4886025119029199 * z, 0.4886025119029199 * x]
    if deg > 1:
        c += [
            1.0925484305920792 * x * y,
            1.0925484305920792 * y * z,
            0.9461746957575601 * zz - 0.31539156525251999,
            1.0925484305920792 * x * z,
            0.5462742152960396 * (xx - yy),
        ]
    if deg > 2:
        c += [
            0.5900435899266435 * y * (3 * xx - yy),
            2.890611442640554 * x * y * z,
            0.4570457994644658 * y * (5 * zz - 1),
            0.3731763325901154 * z * (5 * zz - 3),
            0.4570457994644658 * x * (5 * zz - 1),
            1.445305721320277 * z * (xx - yy),
            0.5900435899266435 * x * (xx - 3 * yy),
        ]
    if deg > 3:
        c += [
            2.5033429417967046 * x * y * (xx - yy),
            1.7701307697799304 * y * z * (3 * xx - yy),
            0.9461746957575601 * x * y * (7 * zz - 1),
            0.6690465435572892 * y * z * (7 * zz - 3),
            0.10578554691520431 * (35 * zz * zz - 30 * zz + 3),
            0.6690465435572892 * x * z * (7 * zz - 3),
            0.47308734787878004 * (xx - yy) * (7 * zz - 1),
            1.7701307697799304 * x * z * (xx - 3 * yy),
            0.6258357354491761 * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)),
        ]
    return torch.stack(c, -1)


def nerf_freq_encode(x: Tensor, num_freq: int, min_exp: float, max_exp: float, include_input: bool) -> Tensor:
    """NeRFEncoding torch path (encodings.py:148-186): sin of [x·2π·2^f] ‖ the same + π/2."""
    freqs = 2 ** torch.linspace(min_exp, max_exp, num_freq)
    s = (2 * torch.pi * x)[..., None] * freqs
    s = s.reshape(*s.shape[:-2], -1)
    enc = torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))
    return torch.cat([enc, x], dim=-1) if include_input else enc


# ----------------------------------------------------------------------------------------
# MLP (a12), trunc_exp (a14), contraction (a9), normalisation/selector (a10)
# ----------------------------------------------------------------------------------------
_ACTS: Dict[str, Callable[[Tensor], Tensor]] = {
    "none": lambda t: t,
    "relu": torch.relu,
    "sigmoid": torch.sigmoid,
    "softplus": torch.nn.functional.softplus,
    "tanh": torch.tanh,
}


def mlp_forward(
    x: Tensor,
    weights: Sequence[Tensor],
    biases: Sequence[Optional[Tensor]],
    skip: Sequence[int] = (),
    act: str = "relu",
    out_act: str = "none",
) -> Tensor:
    """field_components/mlp.py:160-179.  weights[i] is [out_i, in_i] (nn.Linear layout)."""
    h = x
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        if i in skip:
            h = torch.cat([x, h], -1)
        h = torch.nn.functional.linear(h, w, b)
        if i < n - 1:
            h = _ACTS[act](h)
    return _ACTS[out_act](h)


class _TruncExp(torch.autograd.Function):
    """field_components/activations.py:28-41: exp fwd, grad uses exp(clamp(x,-15,15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


def contract_linf(p: Tensor) -> Tensor:
    """SceneContraction(order=inf) (field_components/spatial_distortions.py:66-69)."""
    mag = torch.linalg.norm(p, ord=float("inf"), dim=-1)[..., None]
    return torch.where(mag < 1, p, (2 - (1 / mag)) * (p / mag))


def frustum_positions(origins: Tensor, directions: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """Frustums.get_positions (cameras/rays.py:50-59). origins/directions [R,3], starts/ends [R,S]."""
    return origins[:, None, :] + directions[:, None, :] * (starts[..., None] + ends[..., None]) / 2


def normalize_and_select(p: Tensor, aabb: Tensor, contraction: bool) -> Tuple[Tensor, Tensor]:
    """nerfacto_field.py:205-213 / density_fields.py:95-102 / data/scene_box.py:62-71."""
    if contraction:
        p = contract_linf(p)
        p = (p + 2.0) / 4.0
    else:
        p = (p - aabb[0]) / (aabb[1] - aabb[0])
    sel = ((p > 0.0) & (p < 1.0)).all(dim=-1)
    return p * sel[..., None], sel


# ----------------------------------------------------------------------------------------
# fields (a15, a16, a19, a29)
# ----------------------------------------------------------------------------------------
def density_field(positions: Tensor, P: dict, aabb: Tensor, contraction: bool, avg_init: float) -> Tensor:
    """HashMLPDensityField.get_density (fields/density_fields.py:94-117).  positions [...,3] -> [...,1]."""
    x, sel = normalize_and_select(positions, aabb, contraction)
    enc = hash_encode(x.reshape(-1, 3), P["table"], P["scalings"], P["log2_T"])
    h = mlp_forward(enc, P["w"], P["b"]).view(*positions.shape[:-1], 1)
    return avg_init * trunc_exp(h) * sel[..., None]


def nerfacto_field(
    positions: Tensor,
    directions: Tensor,
    camera_indices: Optional[Tensor],
    P: dict,
    aabb: Tensor,
    contraction: bool,
    avg_init: float,
    training: bool = True,
    use_average_appearance: bool = False,
) -> Tuple[Tensor, Tensor]:
    """NerfactoField.get_density + get_outputs (fields/nerfacto_field.py:203-310), default heads only.

    positions/directions [...,3] (directions already broadcast per sample), camera_indices [...]
    Returns (density [...,1], rgb [...,3]).
    """
    shp = positions.shape[:-1]
    x, sel = normalize_and_select(positions, aabb, contraction)
    enc = hash_encode(x.reshape(-1, 3), P["table"], P["scalings"], P["log2_T"])
    h = mlp_forward(enc, P["w_base"], P["b_base"])
    geo = h.shape[-1] - 1
    dens_pre, feat = torch.split(h, [1, geo], dim=-1)
    density = avg_init * trunc_exp(dens_pre.view(*shp, 1)) * sel[..., None]
    with torch.no_grad():
        sh = sh_components(4, ((directions + 1.0) / 2.0).reshape(-1, 3))
    parts = [sh, feat]
    emb = P.get("embedding")
    if emb is not None:
        if training:
            parts.append(emb[camera_indices.reshape(-1)])
        elif use_average_appearance:
            parts.append(torch.ones(sh.shape[0], emb.shape[1]) * emb.mean(0))
        else:
            parts.append(torch.zeros(sh.shape[0], emb.shape[1]))
    rgb = mlp_forward(torch.cat(parts, -1), P["w_head"], P["b_head"], out_act="sigmoid").view(*shp, 3)
    return density, rgb


def vanilla_nerf_field(positions: Tensor, directions: Tensor, P: dict) -> Tuple[Tensor, Tensor]:
    """NeRFField (fields/vanilla_nerf_field.py:45-107) with the vanilla-nerf encodings
    (models/vanilla_nerf.py:86-92): pos 10 freqs max_exp 8, dir 4 freqs max_exp 4, include_input."""
    shp = positions.shape[:-1]
    pe = nerf_freq_encode(positions.reshape(-1, 3), 10, 0.0, 8.0, True)
    de = nerf_freq_encode(directions.reshape(-1, 3), 4, 0.0, 4.0, True)
    base = mlp_forward(pe, P["w_base"], P["b_base"], skip=P.get("skip", (4,)), out_act="relu")
    density = torch.nn.functional.softplus(torch.nn.functional.linear(base, P["w_sigma"], P["b_sigma"]))
    hh = mlp_forward(torch.cat([de, base], -1), P["w_head"], P["b_head"], out_act="relu")
    rgb = torch.sigmoid(torch.nn.functional.linear(hh, P["w_rgb"], P["b_rgb"]))
    return density.view(*shp, 1), rgb.view(*shp, 3)


# ----------------------------------------------------------------------------------------
# samplers (a5, a6, a7)
# ----------------------------------------------------------------------------------------
def _piecewise_fn(x):
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def _piecewise_inv(x):
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


SPACING = {
    "uniform": (lambda x: x, lambda x: x),
    "piecewise": (_piecewise_fn, _piecewise_inv),
    "lindisp": (lambda x: 1 / x, lambda x: 1 / x),
    "sqrt": (torch.sqrt, lambda x: x ** 2),
    "log": (torch.log, torch.exp),
}


def spacing_to_euclid(bins: Tensor, nears: Tensor, fars: Tensor, kind: str) -> Tensor:
    fn, inv = SPACING[kind]
    s_near, s_far = fn(nears), fn(fars)
    return inv(bins * s_far + (1 - bins) * s_near)


def spaced_sample(
    nears: Tensor, fars: Tensor, num_samples: int, kind: str, jitter: Optional[Tensor]
) -> Tuple[Tensor, Tensor]:
    """SpacedSampler.generate_ray_samples (model_components/ray_samplers.py:78-128).

    nears/fars [R,1]; jitter None (eval), [R,1] (single_jitter) or [R,S+1] uniform(0,1) draws.
    Returns (spacing bins [R,S+1], euclidean bins [R,S+1]).
    """
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, :]
    if jitter is not None:
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * jitter
    else:
        bins = bins.expand(nears.shape[0], -1)
    return bins, spacing_to_euclid(bins, nears, fars, kind)


def pdf_sample(
    existing_bins: Tensor,
    weights: Tensor,
    num_samples: int,
    jitter: Optional[Tensor],
    histogram_padding: float = 0.01,
    eps: float = 1e-5,
    include_original: bool = False,
) -> Dict[str, Tensor]:
    """PDFSampler.generate_ray_samples (model_components/ray_samplers.py:276-372), spacing domain.

    existing_bins [R,S+1], weights [R,S]; jitter None (eval), [R,1] or [R,num_samples+1] in U(0,1).
    Returns dict(bins [R,nb(+S+1)], inds int64 [R,nb], cdf [R,S+1], u [R,nb]).
    """
    nb = num_samples + 1
    w = weights + histogram_padding
    w_sum = torch.sum(w, dim=-1, keepdim=True)
    pad = torch.relu(eps - w_sum)
    w = w + pad / w.shape[-1]
    w_sum = w_sum + pad
    pdf = w / w_sum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb)
    if jitter is not None:
        u = u.expand(cdf.shape[0], nb) + jitter / nb
    else:
        u = (u + 1.0 / (2 * nb)).expand(cdf.shape[0], nb)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, side="right")
    hi = existing_bins.shape[-1] - 1
    below = torch.clamp(inds - 1, 0, hi)
    above = torch.clamp(inds, 0, hi)
    c0, b0 = torch.gather(cdf, -1, below), torch.gather(existing_bins, -1, below)
    c1, b1 = torch.gather(cdf, -1, above), torch.gather(existing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    bins = b0 + t * (b1 - b0)
    if include_original:
        bins, _ = torch.sort(torch.cat([existing_bins, bins], -1), -1)
    return dict(bins=bins.detach(), inds=inds, cdf=cdf, u=u)


# ----------------------------------------------------------------------------------------
# weights + renderers (a21, a22, a23)
# ----------------------------------------------------------------------------------------
def get_weights(deltas: Tensor, density: Tensor) -> Tensor:
    """RaySamples.get_weights (cameras/rays.py:129-152).  deltas, density [R,S,1]."""
    dd = deltas * density
    alphas = 1 - torch.exp(-dd)
    trans = torch.cumsum(dd[..., :-1, :], dim=-2)
    trans = torch.cat([torch.zeros((*trans.shape[:1], 1, 1)), trans], dim=-2)
    return torch.nan_to_num(alphas * torch.exp(-trans))


def composite_rgb(rgb: Tensor, weights: Tensor, background: str = "last_sample", training: bool = True) -> Tensor:
    """RGBRenderer.forward/combine_rgb (model_components/renderers.py:71-119,201-232). rgb [R,S,3], w [R,S,1]."""
    if not training:
        rgb = torch.nan_to_num(rgb)
    comp = torch.sum(weights * rgb, dim=-2)
    acc = torch.sum(weights, dim=-2)
    if background == "last_sample":
        comp = comp + rgb[..., -1, :] * (1.0 - acc)
    elif background == "white":
        comp = comp + torch.ones(3) * (1.0 - acc)
    elif background == "black":
        comp = comp + torch.zeros(3) * (1.0 - acc)
    elif background != "random":
        raise ValueError(background)
    if not training:
        comp = comp.clamp(0.0, 1.0)
    return comp


def accumulation(weights: Tensor) -> Tensor:
    """AccumulationRenderer (renderers.py:292-317)."""
    return torch.sum(weights, dim=-2)


def depth_median(weights: Tensor, starts: Tensor, ends: Tensor) -> Tuple[Tensor, Tensor]:
    """DepthRenderer('median') (renderers.py:354-363). weights/starts/ends [R,S,1] -> (depth [R,1], idx [R,1])."""
    steps = (starts + ends) / 2
    cw = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1)) * 0.5
    idx = torch.searchsorted(cw, split, side="left")
    idx = torch.clamp(idx, 0, steps.shape[-2] - 1)
    return torch.gather(steps[..., 0], dim=-1, index=idx), idx


def depth_expected(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """DepthRenderer('expected') (renderers.py:364-383)."""
    steps = (starts + ends) / 2
    d = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)
    return torch.clip(d, steps.min(), steps.max())


# ----------------------------------------------------------------------------------------
# proposal losses (a24)
# ----------------------------------------------------------------------------------------
def outer_envelope(t0_s: Tensor, t0_e: Tensor, t1_s: Tensor, t1_e: Tensor, y1: Tensor) -> Tensor:
    """losses.py:53-82."""
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    lo = torch.searchsorted(t1_s.contiguous(), t0_s.contiguous(), side="right") - 1
    lo = torch.clamp(lo, min=0, max=y1.shape[-1] - 1)
    hi = torch.searchsorted(t1_e.contiguous(), t0_e.contiguous(), side="right")
    hi = torch.clamp(hi, min=0, max=y1.shape[-1] - 1)
    return torch.take_along_dim(cy1[..., 1:], hi, dim=-1) - torch.take_along_dim(cy1[..., :-1], lo, dim=-1)


def lossfun_outer(t: Tensor, w: Tensor, t_env: Tensor, w_env: Tensor) -> Tensor:
    """losses.py:85-102 (EPS = 1e-7, losses.py:41)."""
    w_outer = outer_envelope(t[..., :-1], t[..., 1:], t_env[..., :-1], t_env[..., 1:], w_env)
    return torch.clip(w - w_outer, min=0) ** 2 / (w + 1.0e-7)


def interlevel_loss(weights_list: List[Tensor], sdist_list: List[Tensor]) -> Tensor:
    """losses.py:113-132.  weights_list[i] [R,S_i], sdist_list[i] [R,S_i+1] (spacing-domain bin edges)."""
    c = sdist_list[-1].detach()
    w = weights_list[-1].detach()
    total = 0.0
    for sd, wp in zip(sdist_list[:-1], weights_list[:-1]):
        total = total + torch.mean(lossfun_outer(c, w, sd, wp))
    return total


def lossfun_distortion(t: Tensor, w: Tensor) -> Tensor:
    """losses.py:135-146."""
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    intra = torch.sum(w ** 2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return inter + intra


def distortion_loss(weights_last: Tensor, sdist_last: Tensor) -> Tensor:
    """losses.py:149-154."""
    return torch.mean(lossfun_distortion(sdist_last, weights_last))


# ----------------------------------------------------------------------------------------
# ray generation (a1, a2) and colliders (a3)
# ----------------------------------------------------------------------------------------
def _undistort(coords: Tensor, dist: Tensor, eps: float = 1e-3, iters: int = 10) -> Tensor:
    """camera_utils.py:375-478 — 10 Newton steps on the OpenCV radial(k1..k4)+tangential(p1,p2) model."""
    k1, k2, k3, k4, p1, p2 = (dist[..., i] for i in range(6))
    xd, yd = coords[..., 0], coords[..., 1]
    x, y = xd, yd
    for _ in range(iters):
        r = x * x + y * y
        d = 1.0 + r * (k1 + r * (k2 + r * (k3 + r * k4)))
        fx = d * x + 2 * p1 * x * y + p2 * (r + 2 * x * x) - xd
        fy = d * y + 2 * p2 * x * y + p1 * (r + 2 * y * y) - yd
        d_r = k1 + r * (2.0 * k2 + r * (3.0 * k3 + r * 4.0 * k4))
        d_x, d_y = 2.0 * x * d_r, 2.0 * y * d_r
        fx_x = d + d_x * x + 2.0 * p1 * y + 6.0 * p2 * x
        fx_y = d_y * x + 2.0 * p1 * x + 2.0 * p2 * y
        fy_x = d_x * y + 2.0 * p2 * y + 2.0 * p1 * x
        fy_y = d + d_y * y + 2.0 * p2 * x + 6.0 * p1 * y
        den = fy_x * fx_y - fx_x * fy_y
        ok = torch.abs(den) > eps
        x = x + torch.where(ok, (fx * fy_y - fy * fx_y) / den, torch.zeros_like(den))
        y = y + torch.where(ok, (fy * fx_x - fx * fy_x) / den, torch.zeros_like(den))
    return torch.stack([x, y], dim=-1)


_NORM_EPS = 4 * 2.220446049250313e-16  # camera_utils.py:28


def generate_rays_perspective(
    c2w: Tensor, fx: Tensor, fy: Tensor, cx: Tensor, cy: Tensor, dist: Optional[Tensor], ray_indices: Tensor,
    coords: Optional[Tensor] = None,
) -> Dict[str, Tensor]:
    """RayGenerator.forward + Cameras._generate_rays_from_coords for PERSPECTIVE cameras
    (model_components/ray_generators.py:41-56; cameras/cameras.py:599-929).

    c2w [C,3,4]; fx,fy,cx,cy [C]; dist [C,6] or None; ray_indices int64 [R,3] = (camera,row,col).
    coords default to pixel centres (row+0.5, col+0.5) (cameras.py:291-319 get_image_coords).
    """
    cam = ray_indices[:, 0]
    if coords is None:
        coords = ray_indices[:, 1:3].float() + 0.5
    y, x = coords[:, 0], coords[:, 1]
    fxr, fyr, cxr, cyr = fx[cam], fy[cam], cx[cam], cy[cam]
    c0 = torch.stack([(x - cxr) / fxr, (y - cyr) / fyr], -1)
    c1 = torch.stack([(x - cxr + 1) / fxr, (y - cyr) / fyr], -1)
    c2 = torch.stack([(x - cxr) / fxr, (y - cyr + 1) / fyr], -1)
    stack = torch.stack([c0, c1, c2], 0)
    if dist is not None and bool((dist[cam] != 0).any()):
        stack = _undistort(stack, dist[cam][None].expand(3, -1, -1))
    stack = stack.clone()
    stack[..., 1] *= -1
    dirs = torch.stack([stack[..., 0], stack[..., 1], -torch.ones_like(stack[..., 0])], -1)
    rot = c2w[cam][:, :3, :3]
    dirs = torch.sum(dirs[..., None, :] * rot, dim=-1)
    norm = torch.maximum(torch.linalg.vector_norm(dirs, dim=-1, keepdim=True), torch.tensor([_NORM_EPS]).to(dirs))
    dirs = dirs / norm
    d0 = dirs[0]
    dx = torch.sqrt(torch.sum((d0 - dirs[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((d0 - dirs[2]) ** 2, dim=-1))
    return dict(
        origins=c2w[cam][:, :3, 3], directions=d0, pixel_area=(dx * dy)[:, None],
        camera_indices=cam[:, None], directions_norm=norm[0],
    )


def aabb_collider(origins: Tensor, directions: Tensor, aabb: Tensor, near_plane: float) -> Tuple[Tensor, Tensor]:
    """AABBBoxCollider (model_components/scene_colliders.py:47-108)."""
    inv = 1.0 / (directions + 1e-6)
    t1 = (aabb[0] - origins) * inv
    t2 = (aabb[1] - origins) * inv
    nears = torch.max(torch.minimum(t1, t2), dim=1).values
    fars = torch.min(torch.maximum(t1, t2), dim=1).values
    nears = torch.clamp(nears, min=near_plane)
    fars = torch.maximum(fars, nears + 1e-6)
    return nears[:, None], fars[:, None]


# ----------------------------------------------------------------------------------------
# packed path: nerfacc 0.5.2 semantics as used by instant-ngp (a25-a27).  PARITY UNPINNED.
# ----------------------------------------------------------------------------------------
def pack_info(ray_indices: Tensor, n_rays: int) -> Tensor:
    """nerfacc.pack_info: per ray (start, count) of its run in the sorted `ray_indices`."""
    counts = torch.bincount(ray_indices, minlength=n_rays)
    starts = torch.cumsum(counts, 0) - counts
    return torch.stack([starts, counts], -1)


def packed_weights(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, ray_indices: Tensor, n_rays: int):
    """nerfacc.render_weight_from_density: alpha=1-exp(-σδ), T=exp(-exclusive_sum_per_ray(σδ)), w=Tα."""
    sd = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sd)
    info = pack_info(ray_indices, n_rays)
    cs = torch.cumsum(sd.double(), 0)  # fp64 so that the per-ray prefix keeps fp32 accuracy on long packs
    excl = cs - sd.double()
    base = torch.zeros(n_rays, dtype=torch.float64)
    nonempty = info[:, 1] > 0
    base[nonempty] = excl[info[nonempty, 0]]
    trans = torch.exp(-(excl - base[ray_indices]).to(sd.dtype))
    return trans * alphas, trans, alphas


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor], ray_indices: Tensor, n_rays: int) -> Tensor:
    """nerfacc.accumulate_along_rays: out[ray] += w * v (index_add)."""
    src = weights[:, None] if values is None else weights[:, None] * values
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype)
    return out.index_add(0, ray_indices, src)


def ray_aabb_intersect(o: Tensor, d: Tensor, aabb: Tensor, near: float = 0.0, far: float = 1e10):
    """Slab test in the style of utils/math.py:138-175 (division by d, no epsilon)."""
    inv = 1.0 / d
    t1 = (aabb[:3] - o) * inv
    t2 = (aabb[3:] - o) * inv
    tmin = torch.max(torch.minimum(t1, t2), dim=-1).values.clamp(min=near)
    tmax = torch.min(torch.maximum(t1, t2), dim=-1).values.clamp(max=far)
    hit = tmax > tmin
    return tmin, tmax, hit


def occgrid_march(
    o: Tensor, d: Tensor, binaries: Tensor, aabb: Tensor, step: float, near: float, far: float,
    cone_angle: float = 0.0, jitter: Optional[Tensor] = None, max_samples_per_ray: int = 1 << 20,
    t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None,
):
    """Occupancy-grid ray marching with nerfacc-style multi-level grids (pure-python; small cases only).

    binaries bool [levels,res,res,res]; level i covers the roi aabb scaled ×2^i about its centre.
    A candidate interval [t, t+dt) (dt = max(t*cone_angle, step)) is emitted when the finest level that
    contains its midpoint has the cell bit set.  Our CUDA traversal implements exactly this rule, so
    indices/counts are bit-exact against this function; against nerfacc itself it is unpinned.
    Returns (ray_indices int64 [M], t_starts [M], t_ends [M]).
    """
    levels, res = binaries.shape[0], binaries.shape[1]
    centre = (aabb[:3] + aabb[3:]) / 2
    half = (aabb[3:] - aabb[:3]) / 2
    big = torch.cat([centre - half * 2 ** (levels - 1), centre + half * 2 ** (levels - 1)])
    tmin, tmax, hit = ray_aabb_intersect(o, d, big, near, far)
    if t_min is not None:
        tmin = torch.maximum(tmin, t_min.reshape(-1))
    if t_max is not None:
        tmax = torch.minimum(tmax, t_max.reshape(-1))
    hit = tmax > tmin
    ri, ts, te = [], [], []
    f32 = torch.float32
    for r in range(o.shape[0]):
        if not bool(hit[r]):
            continue
        t = tmin[r].to(f32)
        if jitter is not None:
            t = t + jitter[r].to(f32) * torch.tensor(step, dtype=f32)
        n = 0
        while bool(t < tmax[r]) and n < max_samples_per_ray:
            dt = torch.maximum(t * torch.tensor(cone_angle, dtype=f32), torch.tensor(step, dtype=f32))
            t1 = t + dt
            mid = (t + t1) * 0.5
            p = o[r] + d[r] * mid
            rel = torch.abs(p - centre) / half
            m = float(rel.max())
            lvl = 0
            while lvl < levels and m > float(1 << lvl):  # finest level whose box contains the point
                lvl += 1
            if lvl < levels:
                scale = 2.0 ** lvl
                q = ((p - centre) / (half * scale) + 1.0) * 0.5
                cell = torch.clamp((q * res).floor().long(), 0, res - 1)
                if bool(binaries[lvl, cell[0], cell[1], cell[2]]):
                    ri.append(r), ts.append(t.clone()), te.append(t1.clone())
                    n += 1
            t = t1
    if not ri:
        return torch.zeros(0, dtype=torch.int64), torch.zeros(0), torch.zeros(0)
    return torch.tensor(ri, dtype=torch.int64), torch.stack(ts), torch.stack(te)


def packed_visibility_prune(ray_indices: Tensor, t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, n_rays: int,
                            early_stop_eps: float, alpha_thre: float, occs_mean: Optional[float] = None):
    """Pruning inside nerfacc's OccGridEstimator.sampling when a sigma_fn is given (training; call site
    model_components/ray_samplers.py:481-493): alpha_thre = min(alpha_thre, occs.mean()); keep samples with
    T >= early_stop_eps and alpha >= alpha_thre (render_visibility_from_density).  PARITY UNPINNED (restated)."""
    if occs_mean is not None:
        alpha_thre = min(alpha_thre, occs_mean)
    _, trans, alphas = packed_weights(t_starts, t_ends, sigmas, ray_indices, n_rays)
    keep = (trans >= early_stop_eps) & (alphas >= alpha_thre)
    return ray_indices[keep], t_starts[keep], t_ends[keep], keep


def occgrid_cell_points(cell_ids: Tensor, jitter: Tensor, res: int, level_aabb: Tensor) -> Tensor:
    """Jittered sample position of grid cells (nerfacc OccGridEstimator._update): coords from the ij-meshgrid flattening
    id = (x*res + y)*res + z; x = lo + ((coord + jitter)/res) * (hi - lo)."""
    coords = torch.stack([cell_ids // (res * res), (cell_ids // res) % res, cell_ids % res], -1).to(torch.float32)
    u = (coords + jitter) / res
    return level_aabb[:3] + u * (level_aabb[3:] - level_aabb[:3])


def occgrid_update(occs: Tensor, levels: int, res: int, aabbs: Tensor, cells: List[Optional[Tensor]], jitters: List[Tensor],
                   occ_eval_fn, occ_thre: float = 1e-2, ema_decay: float = 0.95):
    """One occupancy update (nerfacc OccGridEstimator._update via update_every_n_steps; called by
    models/instant_ngp.py:149-164 with occ_eval_fn = density_fn(x) * render_step_size).  `cells[l]` = sampled cell ids of
    level l (None = all cells: the warm-up case), `jitters[l]` = their [n,3] U(0,1) offsets — the recorded random stream.
    EMA: occs[cell] = max(occs[cell]*decay, occ) with candidates formed from the OLD values; nerfacc's indexed assignment
    leaves the winner among duplicate cells unspecified — restated as the largest candidate.  Threshold
    min(mean(occs), occ_thre) with the mean accumulated in fp64.  Returns (occs, binaries bool [levels*res^3], thre).
    PARITY UNPINNED (nerfacc 0.5.2 sources unavailable)."""
    occs = occs.clone()
    per = res ** 3
    for lvl in range(levels):
        ids = torch.arange(per) if cells[lvl] is None else cells[lvl]
        x = occgrid_cell_points(ids, jitters[lvl], res, aabbs[lvl])
        occ = occ_eval_fn(x).reshape(-1).to(torch.float32).clamp_min(0.0)
        cand = torch.maximum(occs[lvl * per + ids] * ema_decay, occ)
        occs[lvl * per + ids] = 0.0
        occs.scatter_reduce_(0, lvl * per + ids, cand, reduce="amax", include_self=True)
    mean = (occs.double().sum() / occs.numel()).to(torch.float32)
    thre = torch.minimum(mean, torch.tensor(occ_thre, dtype=torch.float32))
    return occs, occs > thre, thre


# ----------------------------------------------------------------------------------------
# full nerfacto training step (a7 + a15..a24), the unit `bench.py --impl reference` times
# ----------------------------------------------------------------------------------------
def nerfacto_forward(
    P: dict, rays: dict, cfg: dict, rng: dict, training: bool = True
) -> Dict[str, Tensor]:
    """ProposalNetworkSampler loop + NerfactoField + renderers + losses
    (ray_samplers.py:576-617; models/nerfacto.py:298-391).

    P: {"props": [density-field params...], "field": nerfacto-field params}
    rays: origins, directions [R,3], nears, fars [R,1], camera_indices [R], (optional) gt rgb [R,3]
    cfg: num_prop_samples (tuple), num_nerf_samples, aabb [2,3], contraction, avg_init, anneal,
         interlevel_mult, distortion_mult, background
    rng: "jitter0" [R,1] for the initial sampler, "jitter_pdf" list of [R,1] per PDF level (None = eval)
    """
    o, d = rays["origins"], rays["directions"]
    nears, fars = rays["nears"], rays["fars"]
    aabb, contraction, avg = cfg["aabb"], cfg["contraction"], cfg["avg_init"]
    n_prop = len(cfg["num_prop_samples"])
    weights_list, sdist_list, eu_list = [], [], []
    sbins = ebins = weights = None
    for lvl in range(n_prop + 1):
        S = cfg["num_prop_samples"][lvl] if lvl < n_prop else cfg["num_nerf_samples"]
        if lvl == 0:
            sbins, ebins = spaced_sample(nears, fars, S, cfg.get("initial_sampler", "piecewise"), rng.get("jitter0"))
        else:
            annealed = torch.pow(weights, cfg.get("anneal", 1.0))
            jit = rng["jitter_pdf"][lvl - 1] if rng.get("jitter_pdf") is not None else None
            sbins = pdf_sample(sbins, annealed, S, jit)["bins"]
            ebins = spacing_to_euclid(sbins, nears, fars, cfg.get("initial_sampler", "piecewise"))
        starts, ends = ebins[:, :-1], ebins[:, 1:]
        if lvl < n_prop:
            pos = frustum_positions(o, d, starts, ends)
            dens = density_field(pos, P["props"][lvl], aabb, contraction, avg)
            weights = get_weights((ends - starts)[..., None], dens)[..., 0]
            weights_list.append(weights)
            sdist_list.append(sbins)
            eu_list.append(ebins)
    starts, ends = ebins[:, :-1], ebins[:, 1:]
    pos = frustum_positions(o, d, starts, ends)
    S = starts.shape[1]
    dirs = d[:, None, :].expand(-1, S, -1)
    cams = rays["camera_indices"][:, None].expand(-1, S)
    dens, rgb = nerfacto_field(pos, dirs, cams, P["field"], aabb, contraction, avg, training=training)
    w = get_weights((ends - starts)[..., None], dens)
    weights_list.append(w[..., 0])
    sdist_list.append(sbins)
    eu_list.append(ebins)
    out = dict(
        rgb=composite_rgb(rgb, w, cfg.get("background", "last_sample"), training),
        accumulation=accumulation(w),
        depth=depth_median(w.detach(), starts[..., None], ends[..., None])[0],
        expected_depth=depth_expected(w, starts[..., None], ends[..., None]),
        weights_list=weights_list, sdist_list=sdist_list, euclid_list=eu_list,
        density=dens, rgb_samples=rgb,
    )
    if "rgb" in rays:
        out["rgb_loss"] = torch.nn.functional.mse_loss(rays["rgb"], out["rgb"])
        if training:
            out["interlevel_loss"] = cfg.get("interlevel_mult", 1.0) * interlevel_loss(weights_list, sdist_list)
            out["distortion_loss"] = cfg.get("distortion_mult", 0.002) * distortion_loss(weights_list[-1], sdist_list[-1])
            out["loss"] = out["rgb_loss"] + out["interlevel_loss"] + out["distortion_loss"]
        else:
            out["loss"] = out["rgb_loss"]
    return out


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-15):
    """torch.optim.Adam (no weight decay, no amsgrad) single-tensor update, in place; step is 1-based."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ----------------------------------------------------------------------------------------------------------
# a4  CameraOptimizer (SO3xR3): nerfstudio/cameras/lie_groups.py:25-58, cameras/camera_optimizers.py:112-153,177-184
# ----------------------------------------------------------------------------------------------------------
def exp_map_so3xr3(tangent: Tensor) -> Tensor:
    """[B,6] (translation | so(3) log-rotation) -> [B,3,4] = [R | t]; |w|^2 is clamped at 1e-4 before the sqrt."""
    w = tangent[:, 3:]
    n = (w * w).sum(1)
    theta = torch.clamp(n, 1e-4).sqrt()
    inv = 1.0 / theta
    f1 = inv * theta.sin()
    f2 = inv * inv * (1.0 - theta.cos())
    K = torch.zeros(w.shape[0], 3, 3, dtype=w.dtype)
    K[:, 0, 1], K[:, 0, 2] = -w[:, 2], w[:, 1]
    K[:, 1, 0], K[:, 1, 2] = w[:, 2], -w[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -w[:, 1], w[:, 0]
    out = torch.zeros(w.shape[0], 3, 4, dtype=w.dtype)
    out[:, :3, :3] = f1[:, None, None] * K + f2[:, None, None] * torch.bmm(K, K) + torch.eye(3, dtype=w.dtype)[None]
    out[:, :3, 3] = tangent[:, :3]
    return out


def camera_opt_apply(pose_adjustment: Tensor, camera_indices: Tensor, origins: Tensor, directions: Tensor):
    """apply_to_raybundle: origins + t[cam], R[cam] @ directions."""
    m = exp_map_so3xr3(pose_adjustment[camera_indices.reshape(-1)])
    return origins + m[:, :3, 3], torch.bmm(m[:, :3, :3], directions[..., None]).squeeze(-1)


def camera_opt_regularizer(pose_adjustment: Tensor, trans_l2_penalty: float = 1e-2, rot_l2_penalty: float = 1e-3) -> Tensor:
    return (pose_adjustment[:, :3].norm(dim=-1).mean() * trans_l2_penalty
            + pose_adjustment[:, 3:].norm(dim=-1).mean() * rot_l2_penalty)


def philox_uniform(n: int, seed: int, draw: int):
    """Uniform [0,1) floats of b2n_step_begin (include/b200nerf.h): element 4q+j = word j of Philox-4x32-10 (Salmon et
    al., SC'11, "Parallel random numbers: as easy as 1, 2, 3") with counter (q_lo, q_hi, draw_lo, draw_hi) and key
    (seed_lo, seed_hi), top 24 bits scaled by 2^-24.  Restated from the paper; the reference draws with torch.rand
    (ray_samplers.py:99-105), whose values are not part of the parity contract."""
    import numpy as np

    q = np.arange((n + 3) // 4, dtype=np.uint64)
    c = [q & 0xFFFFFFFF, q >> np.uint64(32), np.full_like(q, draw & 0xFFFFFFFF), np.full_like(q, (draw >> 32) & 0xFFFFFFFF)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & 0xFFFFFFFF, p1 >> np.uint64(32), p1 & 0xFFFFFFFF
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    words = np.stack(c, axis=1).reshape(-1)[:n]
    return ((words >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
