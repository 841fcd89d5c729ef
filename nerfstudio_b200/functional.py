"""Tensor-level API over the C-ABI: thin wrappers + torch.autograd.Functions.

Every function here launches hand-written sm_100a kernels from libb200nerf.so on the current CUDA stream.
PyTorch is used for device memory, streams and autograd bookkeeping only.  Autocast-safe: the Functions cast
their floating inputs to fp32 (`custom_fwd(cast_inputs=torch.float32)`), as the reference's `trunc_exp`
does (nerfstudio/field_components/activations.py:28-41).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.amp import custom_bwd, custom_fwd

from . import lib
from .lib import B2nGrid, B2nMlp, B2nMlpGrad, call, host_floats, ptr, stream

_fwd = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------
# optional per-kernel timing (CUDA events on the launching stream) for bench.py's roofline line
# ----------------------------------------------------------------------------------------
PROFILE_KERNELS = False
KERNEL_TIMES: dict = {}


class _Timed:
    def __init__(self, key: str, algorithmic_bytes: int):
        self.key, self.bytes = key, algorithmic_bytes

    def __enter__(self):
        if PROFILE_KERNELS:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE_KERNELS:
            self.e1.record()
            KERNEL_TIMES.setdefault(self.key, {"bytes": self.bytes, "events": []})["events"].append((self.e0, self.e1))


def kernel_time_summary() -> dict:
    """key -> {launches, ms_total, ms_avg, bytes_per_launch}; call after torch.cuda.synchronize()."""
    out = {}
    for k, v in KERNEL_TIMES.items():
        ms = [a.elapsed_time(b) for a, b in v["events"]]
        if ms:
            out[k] = {"launches": len(ms), "ms_total": sum(ms), "ms_avg": sum(ms) / len(ms), "bytes_per_launch": v["bytes"]}
    return out


# ----------------------------------------------------------------------------------------
# hash grid
# ----------------------------------------------------------------------------------------
class GridSpec:
    """Host-side description of a multiresolution grid (wraps the C struct; hashable cache of level metadata)."""

    def __init__(self, scales: Sequence[float], log2_hashmap_size: int, n_features: int, mode: str = "torch",
                 resolutions: Optional[Sequence[int]] = None, offsets: Optional[Sequence[int]] = None,
                 sizes: Optional[Sequence[int]] = None, hashed: Optional[Sequence[bool]] = None):
        L = len(scales)
        if not 1 <= L <= lib.MAX_LEVELS:
            raise ValueError(f"num_levels must be in 1..{lib.MAX_LEVELS}")
        if n_features not in (1, 2, 4, 8):
            raise ValueError("features_per_level must be 1, 2, 4 or 8")
        T = 1 << log2_hashmap_size
        g = B2nGrid()
        g.n_levels, g.n_features, g.log2_hashmap_size = L, n_features, log2_hashmap_size
        g.mode = lib.GRID_TORCH if mode == "torch" else lib.GRID_TCNN
        for l in range(L):
            g.scale[l] = float(scales[l])
            g.resolution[l] = int(resolutions[l]) if resolutions is not None else 0
            g.offset[l] = int(offsets[l]) if offsets is not None else l * T
            g.size[l] = int(sizes[l]) if sizes is not None else T
            g.hashed[l] = int(hashed[l]) if hashed is not None else 1
        self.c = g
        self.n_levels, self.n_features, self.mode = L, n_features, mode
        self.out_dim = L * n_features
        self.n_rows = max(int(g.offset[l]) + int(g.size[l]) for l in range(L))

    @staticmethod
    def tcnn(num_levels: int, base_res: int, per_level_scale: float, log2_hashmap_size: int, n_features: int):
        """tiny-cuda-nn HashGrid level table: dense coarse levels, hashed fine ones."""
        scales, res, offs, sizes, hashed, off = [], [], [], [], [], 0
        for l in range(num_levels):
            s = math.exp2(l * math.log2(per_level_scale)) * base_res - 1.0
            r = int(math.ceil(s)) + 1
            n = (r ** 3 + 7) // 8 * 8
            size = min(n, 1 << log2_hashmap_size)
            scales.append(s), res.append(r), offs.append(off), sizes.append(size), hashed.append(r ** 3 > size)
            off += size
        return GridSpec(scales, log2_hashmap_size, n_features, "tcnn", res, offs, sizes, hashed)


def hashgrid_forward(x: Tensor, table: Tensor, grid: GridSpec, want_indices: bool = False):
    """x [N,3] fp32 in [0,1], table [rows,F] -> y [N, L*F] (and int64 [N,L,8] corner rows if asked)."""
    x, table = _c(x), _c(table)
    n = x.shape[0]
    y = torch.empty(n, grid.out_dim, device=x.device, dtype=torch.float32)
    idx = torch.empty(n, grid.n_levels, 8, device=x.device, dtype=torch.int64) if want_indices else None
    # algorithmic bytes: 8 corners x F floats per (point, level)  
    with _Timed(f"hashgrid_fwd[N={n},L={grid.n_levels},F={grid.n_features}]", n * grid.n_levels * 8 * grid.n_features * 4):
        call("b2n_hashgrid_fwd", C.byref(grid.c), ptr(x), ptr(table), n, ptr(y), ptr(idx, torch.int64), stream())
    return (y, idx) if want_indices else y


def hashgrid_backward(x: Tensor, table: Tensor, dy: Tensor, grid: GridSpec, dtable: Optional[Tensor] = None,
                      want_dx: bool = False):
    x, table, dy = _c(x), _c(table), _c(dy)
    if dtable is None:
        dtable = torch.zeros_like(table)
    dx = torch.empty_like(x) if want_dx else None
    n = x.shape[0]
    # scatter-add counted as read + write of every touched row
    with _Timed(f"hashgrid_bwd[N={n},L={grid.n_levels},F={grid.n_features}]", 2 * n * grid.n_levels * 8 * grid.n_features * 4):
        call("b2n_hashgrid_bwd", C.byref(grid.c), ptr(x), ptr(table), ptr(dy), n, ptr(dtable), ptr(dx), stream())
    return dtable, dx


class _HashGridFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, x, table, grid):
        ctx.grid = grid
        ctx.save_for_backward(x, table)
        return hashgrid_forward(x, table, grid)

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, table = ctx.saved_tensors
        dtable, dx = hashgrid_backward(x, table, dy.float(), ctx.grid, want_dx=ctx.needs_input_grad[0])
        return dx, (dtable if ctx.needs_input_grad[1] else None), None


def hash_encode(x: Tensor, table: Tensor, grid: GridSpec) -> Tensor:
    return _HashGridFn.apply(x, table, grid)


# ----------------------------------------------------------------------------------------
# tiny MLP
# ----------------------------------------------------------------------------------------
class MlpSpec:
    def __init__(self, in_dim: int, out_dims: Sequence[int], skip: Sequence[int] = (), hidden_act: str = "relu",
                 out_act: str = "none", bias: bool = True):
        if len(out_dims) > lib.MAX_MLP_LAYERS:
            raise ValueError(f"at most {lib.MAX_MLP_LAYERS} layers")
        self.in_dim, self.out_dims, self.skip = int(in_dim), [int(o) for o in out_dims], set(int(s) for s in skip)
        self.hidden_act, self.out_act, self.bias = hidden_act, out_act, bias
        self.hidden_width = sum(self.out_dims[:-1])

    def layer_in(self, i: int) -> int:
        prev = self.in_dim if i == 0 else self.out_dims[i - 1]
        return prev + self.in_dim if i in self.skip else prev

    def fits_fused_kernel(self) -> bool:
        """Shared-memory budget of the backward kernel (the larger one) — mirrors mlp.cu:bwd_smem."""
        r8 = lambda v: (v + 7) // 8 * 8
        w_total, kmax, wmax = 0, r8(self.in_dim), 8
        for i, out in enumerate(self.out_dims):
            inn = self.layer_in(i)
            w_total += (out * r8(inn) + 3) // 4 * 4 + r8(out)
            kmax, wmax = max(kmax, r8(inn)), max(wmax, r8(out))
        smem = 4 * (2 * w_total + (wmax + 2 * kmax + (self.in_dim if self.skip else 0)) * 129)
        return smem <= 227 * 1024

    def struct(self, weights: Sequence[Tensor], biases: Sequence[Optional[Tensor]]) -> B2nMlp:
        m = B2nMlp()
        m.n_layers, m.in_dim = len(self.out_dims), self.in_dim
        m.hidden_act, m.out_act = lib.ACT[self.hidden_act], lib.ACT[self.out_act]
        for i, out in enumerate(self.out_dims):
            w = weights[i]
            if tuple(w.shape) != (out, self.layer_in(i)):
                raise ValueError(f"layer {i}: weight shape {tuple(w.shape)} != {(out, self.layer_in(i))}")
            m.out_dims[i], m.skip[i] = out, int(i in self.skip)
            m.w[i] = ptr(w).value
            m.b[i] = ptr(biases[i]).value if biases[i] is not None else None
        return m


def mlp_forward(spec: MlpSpec, x: Tensor, weights, biases, save_hidden: bool):
    x = _c(x)
    n = x.shape[0]
    y = torch.empty(n, spec.out_dims[-1], device=x.device, dtype=torch.float32)
    hidden = torch.empty(max(spec.hidden_width, 1) * n, device=x.device, dtype=torch.float32) if save_hidden else None
    m = spec.struct([_c(w) for w in weights], [None if b is None else _c(b) for b in biases])
    call("b2n_mlp_fwd", C.byref(m), ptr(x), n, ptr(y), ptr(hidden), stream())
    return y, hidden


def mlp_backward(spec: MlpSpec, x, y, hidden, dy, weights, biases, dws, dbs, want_dx: bool):
    m = spec.struct(weights, biases)
    g = B2nMlpGrad()
    for i in range(len(spec.out_dims)):
        g.dw[i] = ptr(dws[i]).value if dws[i] is not None else None
        g.db[i] = ptr(dbs[i]).value if dbs[i] is not None else None
    dx = torch.empty_like(x) if want_dx else None
    call("b2n_mlp_bwd", C.byref(m), C.byref(g), ptr(x), ptr(y), ptr(hidden), ptr(_c(dy)), x.shape[0], ptr(dx), stream())
    return dx


def mlp_tc_supported(spec: MlpSpec) -> bool:
    """Eligibility of the tensor-core (tcgen05, 3xTF32) kernels — mirrors mlp_tc.cu:tc_build."""
    return (len(spec.out_dims) <= 4 and spec.in_dim <= 64 and not spec.skip and max(spec.out_dims) <= 64
            and all(o % 4 == 0 for o in spec.out_dims[:-1]) and spec.hidden_act in ("relu", "none")
            and spec.out_act in ("none", "sigmoid", "relu"))


def mlp_tc_pack(spec: MlpSpec, weights, biases) -> Tensor:
    """Packed operand images (forward W and backward W^T: hi/lo split, UMMA layout) of the current weights in a fresh
    128-byte aligned device workspace — what the *_ws kernels stage into shared memory with TMA."""
    m = spec.struct([_c(w) for w in weights], [None if b is None else _c(b) for b in biases])
    nbytes = int(lib.load().b2n_mlp_tc_workspace_bytes(C.byref(m)))
    if nbytes <= 0:
        raise ValueError("network outside the tensor-core MLP's shape range")
    buf = torch.zeros(nbytes // 4 + 64, device=weights[0].device, dtype=torch.float32)
    off = (-buf.data_ptr()) % 128 // 4
    ws = buf[off: off + nbytes // 4]
    call("b2n_mlp_tc_pack", C.byref(m), ptr(ws), stream())
    return ws


def mlp_tc_forward(spec: MlpSpec, x: Tensor, weights, biases, save_hidden: bool, x_stride: Optional[int] = None,
                   workspace: Optional[Tensor] = None):
    """Tensor-core forward.  x [N, >=in_dim] row-major (row stride x_stride floats).  workspace: mlp_tc_pack's image."""
    x = _c(x)
    n = x.shape[0]
    y = torch.empty(n, spec.out_dims[-1], device=x.device, dtype=torch.float32)
    hidden = torch.empty(max(spec.hidden_width, 1) * n, device=x.device, dtype=torch.float32) if save_hidden else None
    m = spec.struct([_c(w) for w in weights], [None if b is None else _c(b) for b in biases])
    call("b2n_mlp_tc_fwd_ws", C.byref(m), ptr(x), x_stride or x.shape[1], n, ptr(y), ptr(hidden), ptr(workspace), stream())
    return y, hidden


def mlp_tc_backward(spec: MlpSpec, x, y, hidden, dy, weights, biases, dws, dbs, want_dx: bool,
                    workspace: Optional[Tensor] = None):
    m = spec.struct(weights, biases)
    g = B2nMlpGrad()
    for i in range(len(spec.out_dims)):
        g.dw[i] = ptr(dws[i]).value if dws[i] is not None else None
        g.db[i] = ptr(dbs[i]).value if dbs[i] is not None else None
    dx = torch.empty(x.shape[0], spec.in_dim, device=x.device, dtype=torch.float32) if want_dx else None
    call("b2n_mlp_tc_bwd_ws", C.byref(m), C.byref(g), ptr(x), x.shape[1], ptr(y), ptr(hidden), ptr(_c(dy)), x.shape[0],
         ptr(dx), spec.in_dim, ptr(workspace), stream())
    return dx


class _MlpFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, spec, x, *params):
        L = len(spec.out_dims)
        weights, biases = list(params[:L]), list(params[L:]) if spec.bias else [None] * L
        weights = [_c(w) for w in weights]
        biases = [None if b is None else _c(b) for b in biases]
        x = _c(x)
        need = any(ctx.needs_input_grad)
        y, hidden = mlp_forward(spec, x, weights, biases, save_hidden=need)
        ctx.spec = spec
        if need:
            ctx.save_for_backward(x, y, hidden, *weights, *[b for b in biases if b is not None])
        return y

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        spec = ctx.spec
        L = len(spec.out_dims)
        x, y, hidden, *rest = ctx.saved_tensors
        weights = rest[:L]
        biases = list(rest[L:]) if spec.bias else [None] * L
        dws = [torch.zeros_like(w) for w in weights]
        dbs = [None if b is None else torch.zeros_like(b) for b in biases]
        dx = mlp_backward(spec, x, y, hidden, dy.float(), weights, biases, dws, dbs, want_dx=ctx.needs_input_grad[1])
        return (None, dx, *dws, *[d for d in dbs if d is not None])


def mlp(spec: MlpSpec, x: Tensor, weights: Sequence[Tensor], biases: Sequence[Optional[Tensor]]) -> Tensor:
    params = list(weights) + ([b for b in biases] if spec.bias else [])
    return _MlpFn.apply(spec, x, *params)


class _LinearFn(torch.autograd.Function):
    """One wide dense layer y = act(x W^T + b) on the tiled fp32 GEMM (b2n_linear_fwd / b2n_linear_bwd)."""

    @staticmethod
    @_fwd
    def forward(ctx, x, w, b, act):
        x, w = _c(x), _c(w)
        n, out = x.shape[0], w.shape[0]
        y = torch.empty(n, out, device=x.device, dtype=torch.float32)
        call("b2n_linear_fwd", ptr(x), n, x.shape[1], x.shape[1], ptr(w), ptr(None if b is None else _c(b)), out, lib.ACT[act],
             ptr(y), stream())
        ctx.act = act
        ctx.save_for_backward(x, w, y)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        n, out = y.shape
        dz = torch.empty_like(y)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        db = torch.zeros(out, device=x.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        call("b2n_linear_bwd", ptr(x), n, x.shape[1], x.shape[1], ptr(w), out, lib.ACT[ctx.act], ptr(y), ptr(_c(dy.float())),
             ptr(dz), ptr(dx), x.shape[1], ptr(dw), ptr(db), stream())
        return dx, dw, db, None


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], act: str = "none") -> Tensor:
    return _LinearFn.apply(x, weight, bias, act)


# ----------------------------------------------------------------------------------------
# encodings
# ----------------------------------------------------------------------------------------
def sh_encode(dirs: Tensor, levels: int, remap01: bool = False) -> Tensor:
    """Real SH basis (no grad, as the reference's @torch.no_grad torch path)."""
    if not 1 <= levels <= 5:
        raise ValueError(f"Spherical harmonic encoding only supports 1 to 5 levels, requested {levels}")
    d = _c(dirs.detach().float())
    out = torch.empty(d.shape[0], levels * levels, device=d.device, dtype=torch.float32)
    call("b2n_sh_fwd", ptr(d), d.shape[0], levels, int(remap01), ptr(out), stream())
    return out


class _FreqFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, x, freqs, include_input):
        x = _c(x)
        n, d = x.shape
        fr = host_floats(freqs)
        width = 2 * d * len(freqs) + (d if include_input else 0)
        out = torch.empty(n, width, device=x.device, dtype=torch.float32)
        call("b2n_freq_fwd", ptr(x), n, d, C.cast(fr, C.c_void_p), len(freqs), int(include_input), ptr(out), stream())
        ctx.freqs, ctx.include_input = freqs, include_input
        ctx.save_for_backward(x)
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        fr = host_floats(ctx.freqs)
        dx = torch.empty_like(x)
        call("b2n_freq_bwd", ptr(x), ptr(_c(dout.float())), x.shape[0], x.shape[1], C.cast(fr, C.c_void_p),
             len(ctx.freqs), int(ctx.include_input), ptr(dx), stream())
        return dx, None, None


def freq_encode(x: Tensor, freqs: Sequence[float], include_input: bool) -> Tensor:
    return _FreqFn.apply(x, tuple(float(f) for f in freqs), bool(include_input))


# ----------------------------------------------------------------------------------------
# positions / density activation
# ----------------------------------------------------------------------------------------
def positions_to_unit_cube(origins: Tensor, directions: Optional[Tensor], ebins: Optional[Tensor],
                           contraction: bool, aabb: Optional[Sequence[float]]) -> Tuple[Tensor, Tensor]:
    """Ray form: origins/directions [R,3] + euclidean bin edges [R,S+1] -> x [R*S,3], selector uint8 [R*S].
    Point form (directions None): origins is positions [N,3]."""
    o = _c(origins.float())
    box = host_floats(aabb) if aabb is not None else None
    boxp = C.cast(box, C.c_void_p) if box is not None else C.c_void_p(0)
    if directions is None:
        n = o.shape[0]
        x = torch.empty(n, 3, device=o.device, dtype=torch.float32)
        sel = torch.empty(n, device=o.device, dtype=torch.uint8)
        call("b2n_positions_fwd", ptr(o), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), 0, n, 1, int(contraction), boxp,
             ptr(x), ptr(sel, torch.uint8), stream())
        return x, sel
    d = _c(directions.float())
    iv = ebins if isinstance(ebins, Intervals) else Intervals.from_edges(ebins)
    R, S = iv.R, iv.S
    x = torch.empty(R * S, 3, device=o.device, dtype=torch.float32)
    sel = torch.empty(R * S, device=o.device, dtype=torch.uint8)
    call("b2n_positions_fwd", ptr(o), ptr(d), iv.p_starts, iv.p_ends, iv.stride, R, S, int(contraction), boxp, ptr(x),
         ptr(sel, torch.uint8), stream())
    return x, sel


class _PositionsFn(torch.autograd.Function):
    """Differentiable ray form of positions_to_unit_cube: gradients w.r.t. x flow to origins / directions (used when the
    rays carry a gradient, i.e. behind CameraOptimizer.apply_to_raybundle)."""

    @staticmethod
    def forward(ctx, origins, directions, iv, contraction, aabb):
        x, sel = positions_to_unit_cube(origins.detach(), directions.detach(), iv, contraction, aabb)
        ctx.iv, ctx.contraction, ctx.aabb = iv, contraction, aabb
        ctx.save_for_backward(origins.detach(), directions.detach())
        ctx.mark_non_differentiable(sel)
        return x, sel

    @staticmethod
    def backward(ctx, dx, _dsel):
        o, d = ctx.saved_tensors
        iv = ctx.iv
        o, d = _c(o.float()), _c(d.float())
        d_o, d_d = torch.empty_like(o), torch.empty_like(d)
        box = host_floats(ctx.aabb) if ctx.aabb is not None else None
        call("b2n_positions_bwd", ptr(o), ptr(d), iv.p_starts, iv.p_ends, iv.stride, iv.R, iv.S, int(ctx.contraction),
             C.cast(box, C.c_void_p) if box is not None else C.c_void_p(0), ptr(_c(dx.float())), 0, ptr(d_o), ptr(d_d), stream())
        return d_o, d_d, None, None, None


def positions_to_unit_cube_diff(origins: Tensor, directions: Tensor, ebins, contraction: bool, aabb) -> Tuple[Tensor, Tensor]:
    iv = ebins if isinstance(ebins, Intervals) else Intervals.from_edges(ebins)
    return _PositionsFn.apply(origins, directions, iv, contraction, aabb)


class _DensityActFn(torch.autograd.Function):
    """density = avg_init * trunc_exp(h) * selector."""

    @staticmethod
    @_fwd
    def forward(ctx, h, sel, avg_init):
        h = _c(h)
        n = h.numel()
        out = torch.empty(n, device=h.device, dtype=torch.float32)
        call("b2n_density_act_fwd", ptr(h), 1, ptr(sel, torch.uint8), n, float(avg_init), ptr(out), stream())
        ctx.avg = float(avg_init)
        ctx.save_for_backward(h, sel)
        return out.view(h.shape)

    @staticmethod
    @_bwd
    def backward(ctx, g):
        h, sel = ctx.saved_tensors
        dh = torch.empty_like(h)
        call("b2n_density_act_bwd", ptr(h), 1, ptr(sel, torch.uint8), ptr(_c(g.float())), h.numel(), ctx.avg, ptr(dh), 1,
             stream())
        return dh, None, None


def density_activation(h: Tensor, sel: Optional[Tensor], avg_init: float) -> Tensor:
    return _DensityActFn.apply(h, sel, avg_init)


# ----------------------------------------------------------------------------------------
# fused proposal density field
# ----------------------------------------------------------------------------------------
def density_field_supported(grid: GridSpec, spec: MlpSpec) -> bool:
    """Shape family of the fused kernel (density_fused.cu:check_shape): F=2, <= 8 levels, MLP in->16->1, ReLU."""
    return (grid.n_features == 2 and grid.n_levels <= 8 and spec.in_dim == 2 * grid.n_levels
            and spec.out_dims == [16, 1] and not spec.skip and spec.hidden_act == "relu" and spec.out_act == "none")


def _geom(origins, directions, iv):
    o = _c(origins.float())
    if directions is None:
        return o, None, C.c_void_p(0), C.c_void_p(0), 0, o.shape[0], 1
    return o, _c(directions.float()), iv.p_starts, iv.p_ends, iv.stride, iv.R, iv.S


def density_field_forward(grid: GridSpec, spec: MlpSpec, table, weights, biases, origins, directions, iv, contraction,
                          aabb, avg_init: float) -> Tensor:
    """One launch: ray samples (or raw positions when directions is None) -> density [R*S]."""
    o, d, ps, pe, stride, R, S = _geom(origins, directions, iv)
    m = spec.struct([_c(w) for w in weights], [None if b is None else _c(b) for b in biases])
    box = host_floats(aabb) if aabb is not None else None
    dens = torch.empty(R * S, device=o.device, dtype=torch.float32)
    call("b2n_density_field_fwd", C.byref(grid.c), C.byref(m), ptr(_c(table)), ptr(o), ptr(d), ps, pe, stride, R, S,
         int(contraction), C.cast(box, C.c_void_p) if box is not None else C.c_void_p(0), float(avg_init), ptr(dens),
         stream())
    return dens


def density_field_backward(grid: GridSpec, spec: MlpSpec, table, weights, biases, origins, directions, iv, contraction,
                           aabb, avg_init: float, d_density: Tensor, dtable: Tensor, dws, dbs, compact: bool = True) -> None:
    o, d, ps, pe, stride, R, S = _geom(origins, directions, iv)
    m = spec.struct(weights, biases)
    g = B2nMlpGrad()
    for i in range(2):
        g.dw[i] = ptr(dws[i]).value if dws[i] is not None else None
        g.db[i] = ptr(dbs[i]).value if dbs[i] is not None else None
    box = host_floats(aabb) if aabb is not None else None
    dd = _c(d_density.float())
    live = torch.empty(R * S + 4, device=dd.device, dtype=torch.int32) if compact else None  # scratch: count + indices
    call("b2n_density_field_bwd_ws", C.byref(grid.c), C.byref(m), C.byref(g), ptr(_c(table)), ptr(o), ptr(d), ps, pe, stride,
         R, S, int(contraction), C.cast(box, C.c_void_p) if box is not None else C.c_void_p(0), float(avg_init),
         ptr(dd), ptr(dtable), ptr(live, torch.int32), stream())


class _DensityFieldFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, meta, table, w1, b1, w2, b2):
        grid, spec, origins, directions, iv, contraction, aabb, avg = meta
        table, w1, b1, w2, b2 = _c(table), _c(w1), _c(b1), _c(w2), _c(b2)
        ctx.meta = meta
        ctx.save_for_backward(table, w1, b1, w2, b2)
        return density_field_forward(grid, spec, table, [w1, w2], [b1, b2], origins, directions, iv, contraction, aabb, avg)

    @staticmethod
    @_bwd
    def backward(ctx, d_density):
        grid, spec, origins, directions, iv, contraction, aabb, avg = ctx.meta
        table, w1, b1, w2, b2 = ctx.saved_tensors
        dtable = torch.zeros_like(table)
        dws, dbs = [torch.zeros_like(w1), torch.zeros_like(w2)], [torch.zeros_like(b1), torch.zeros_like(b2)]
        density_field_backward(grid, spec, table, [w1, w2], [b1, b2], origins, directions, iv, contraction, aabb, avg,
                               d_density, dtable, dws, dbs)
        return None, dtable, dws[0], dbs[0], dws[1], dbs[1]


def density_field(grid, spec, table, weights, biases, origins, directions, iv, contraction, aabb, avg_init) -> Tensor:
    """Differentiable (w.r.t. table and network) fused density field; density [R*S]."""
    meta = (grid, spec, origins.detach(), None if directions is None else directions.detach(), iv, contraction, aabb,
            float(avg_init))
    return _DensityFieldFn.apply(meta, table, weights[0], biases[0], weights[1], biases[1])


# ----------------------------------------------------------------------------------------
# samplers
# ----------------------------------------------------------------------------------------
_LINSPACE_CACHE = {}


def _linspace(start: float, end: float, steps: int, device) -> Tensor:
    """torch.linspace evaluated on the CPU (the oracle's arithmetic) and cached on the device."""
    key = (start, end, steps, str(device))
    if key not in _LINSPACE_CACHE:
        _LINSPACE_CACHE[key] = torch.linspace(start, end, steps).to(device)
    return _LINSPACE_CACHE[key]


def spaced_sample(nears: Tensor, fars: Tensor, num_samples: int, spacing: str, jitter: Optional[Tensor]):
    """-> (spacing bins [R,S+1], euclidean bins [R,S+1]).  jitter None | [R,1] | [R,S+1]."""
    n, f = _c(nears.float().reshape(-1)), _c(fars.float().reshape(-1))
    R = n.shape[0]
    lin = _linspace(0.0, 1.0, num_samples + 1, n.device)
    per_bin = 0
    if jitter is not None:
        jitter = _c(jitter.float())
        per_bin = int(jitter.numel() != R)
        if per_bin and jitter.numel() != R * (num_samples + 1):
            raise ValueError("jitter must be [R,1] or [R,S+1]")
    sb = torch.empty(R, num_samples + 1, device=n.device, dtype=torch.float32)
    eb = torch.empty_like(sb)
    call("b2n_spaced_sample", ptr(n), ptr(f), ptr(lin), ptr(jitter), per_bin, R, num_samples, lib.SPACING[spacing],
         ptr(sb), ptr(eb), stream())
    return sb, eb


def pdf_sample(sbins: Tensor, weights: Tensor, num_samples: int, jitter: Optional[Tensor], nears: Tensor, fars: Tensor,
               spacing: str, anneal: float = 1.0, histogram_padding: float = 0.01, eps: float = 1e-5,
               want_aux: bool = False):
    """Inverse-CDF resampling (include_original=False).  -> (new spacing bins [R,nb], new euclidean bins [R,nb]
    [, cdf [R,S+1], inds int64 [R,nb]])."""
    sb, w = _c(sbins.float()), _c(weights.detach().float())
    R, S = w.shape
    nb = num_samples + 1
    u_base = _linspace(0.0, 1.0 - (1.0 / nb), nb, sb.device)
    per_bin = 0
    if jitter is not None:
        jitter = _c(jitter.float())
        per_bin = int(jitter.numel() != R)
    n, f = _c(nears.float().reshape(-1)), _c(fars.float().reshape(-1))
    new_sb = torch.empty(R, nb, device=sb.device, dtype=torch.float32)
    new_eb = torch.empty_like(new_sb)
    cdf = torch.empty(R, S + 1, device=sb.device, dtype=torch.float32) if want_aux else None
    inds = torch.empty(R, nb, device=sb.device, dtype=torch.int64) if want_aux else None
    call("b2n_pdf_sample", ptr(sb), ptr(w), ptr(u_base), ptr(jitter), per_bin, ptr(n), ptr(f), R, S, nb, float(anneal),
         C.c_void_p(0), float(histogram_padding), float(eps), lib.SPACING[spacing], ptr(new_sb), ptr(new_eb), ptr(cdf),
         ptr(inds, torch.int64), stream())
    return (new_sb, new_eb, cdf, inds) if want_aux else (new_sb, new_eb)


def torch_row_sum(x: Tensor) -> Tensor:
    """Row sums of x [R,S] in the summation order of torch's CPU kernel (bit-identical to x.cpu().sum(-1))."""
    x = _c(x.float())
    out = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    call("b2n_torch_row_sum", ptr(x), x.shape[0], x.shape[1], ptr(out), stream())
    return out


# ----------------------------------------------------------------------------------------
# weights / compositing / losses
# ----------------------------------------------------------------------------------------
class Intervals:
    """Sample intervals along rays as the kernels take them: starts/ends pointers + a row stride.

    `from_edges(e)` wraps an [R,S+1] edge array (starts=e, ends=e+1 element, stride S+1) — what every sampler of
    the reference produces; `from_pairs(starts, ends)` takes two independent [R,S] arrays (packed or hand-built)."""

    def __init__(self, base: Tensor, ends: Optional[Tensor], n_rays: int, n_samples: int, stride: int):
        self.base, self.ends_t, self.R, self.S, self.stride = base, ends, n_rays, n_samples, stride

    @staticmethod
    def from_edges(edges: Tensor) -> "Intervals":
        e = _c(edges.detach().float())
        return Intervals(e, None, e.shape[0], e.shape[1] - 1, e.shape[1])

    @staticmethod
    def from_pairs(starts: Tensor, ends: Tensor) -> "Intervals":
        s, e = _c(starts.detach().float()), _c(ends.detach().float())
        return Intervals(s, e, s.shape[0], s.shape[1], s.shape[1])

    @property
    def p_starts(self):
        return ptr(self.base)

    @property
    def p_ends(self):
        return ptr(self.ends_t) if self.ends_t is not None else C.c_void_p(self.base.data_ptr() + 4)

    def starts(self) -> Tensor:
        return self.base[:, : self.S]

    def ends(self) -> Tensor:
        return self.ends_t if self.ends_t is not None else self.base[:, 1:]


class _WeightsFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, iv, density):
        d = _c(density)
        w = torch.empty_like(d)
        call("b2n_weights_fwd", iv.p_starts, iv.p_ends, iv.stride, ptr(d), iv.R, iv.S, ptr(w), stream())
        ctx.iv = iv
        ctx.save_for_backward(d)
        return w

    @staticmethod
    @_bwd
    def backward(ctx, dw):
        (d,) = ctx.saved_tensors
        iv = ctx.iv
        dd = torch.empty_like(d)
        call("b2n_weights_bwd", iv.p_starts, iv.p_ends, iv.stride, ptr(d), ptr(_c(dw.float())), iv.R, iv.S, ptr(dd),
             stream())
        return None, dd


def _intervals(x) -> Intervals:
    return x if isinstance(x, Intervals) else Intervals.from_edges(x)


def render_weights(ebins, density: Tensor) -> Tensor:
    """ebins: [R,S+1] euclidean edges (or an Intervals), density [R,S] -> weights [R,S]."""
    return _WeightsFn.apply(_intervals(ebins), density)


def _bg_args(background):
    if background is None or (isinstance(background, str) and background == "random"):
        return lib.BG_NONE, C.c_void_p(0), None
    if isinstance(background, str) and background == "last_sample":
        return lib.BG_LAST_SAMPLE, C.c_void_p(0), None
    if isinstance(background, str):
        background = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0)}[background]
    arr = host_floats(background)
    return lib.BG_CONSTANT, C.cast(arr, C.c_void_p), arr


class _CompositeFn(torch.autograd.Function):
    """-> (rgb [R,3], accumulation [R], expected depth [R] (unclipped))."""

    @staticmethod
    @_fwd
    def forward(ctx, rgb, weights, iv, background, eval_mode):
        rgb, w = _c(rgb), _c(weights)
        R, S = w.shape
        mode, bgp, keep = _bg_args(background)
        out = torch.empty(R, 3, device=w.device, dtype=torch.float32)
        acc = torch.empty(R, device=w.device, dtype=torch.float32)
        dep = torch.empty(R, device=w.device, dtype=torch.float32)
        call("b2n_composite_fwd", ptr(rgb), ptr(w), iv.p_starts, iv.p_ends, iv.stride, R, S, mode, bgp, int(eval_mode),
             ptr(out), ptr(acc), ptr(dep), C.c_void_p(0), C.c_void_p(0), stream())
        ctx.background, ctx.iv = background, iv
        ctx.save_for_backward(rgb, w)
        return out, acc, dep

    @staticmethod
    @_bwd
    def backward(ctx, d_out, d_acc, d_dep):
        rgb, w = ctx.saved_tensors
        iv = ctx.iv
        R, S = w.shape
        mode, bgp, keep = _bg_args(ctx.background)
        d_rgb = torch.empty_like(rgb)
        d_w = torch.empty_like(w)
        f = lambda t: None if t is None else _c(t.float())
        call("b2n_composite_bwd", ptr(rgb), ptr(w), iv.p_starts, iv.p_ends, iv.stride, ptr(f(d_out)), ptr(f(d_acc)),
             ptr(f(d_dep)), R, S, mode, bgp, ptr(d_rgb), ptr(d_w), stream())
        return d_rgb, d_w, None, None, None


def composite(rgb: Tensor, weights: Tensor, ebins, background="last_sample", eval_mode: bool = False):
    """rgb [R,S,3], weights [R,S], ebins [R,S+1] | Intervals -> (rgb [R,3], acc [R], expected depth [R] pre-clip)."""
    return _CompositeFn.apply(rgb, weights, _intervals(ebins), background, eval_mode)


def accumulate(weights: Tensor) -> Tensor:
    """AccumulationRenderer on dense samples: [R,S] -> [R] (differentiable through torch: a plain row sum)."""
    return weights.sum(dim=-1)


def median_depth(weights: Tensor, ebins) -> Tuple[Tensor, Tensor]:
    w = _c(weights.detach().float())
    iv = _intervals(ebins)
    R, S = w.shape
    dep = torch.empty(R, device=w.device, dtype=torch.float32)
    idx = torch.empty(R, device=w.device, dtype=torch.int64)
    call("b2n_composite_fwd", C.c_void_p(0), ptr(w), iv.p_starts, iv.p_ends, iv.stride, R, S, lib.BG_NONE, C.c_void_p(0),
         0, C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), ptr(dep), ptr(idx, torch.int64), stream())
    return dep, idx


class _InterlevelFn(torch.autograd.Function):
    """sum over rows of the proposal-envelope loss for one proposal level; grad only w.r.t. wp."""

    @staticmethod
    @_fwd
    def forward(ctx, c, w, cp, wp):
        c, w, cp, wp = _c(c), _c(w), _c(cp), _c(wp)
        R, Sc, Sp = w.shape[0], w.shape[1], wp.shape[1]
        rows = torch.empty(R, device=w.device, dtype=torch.float32)
        d_wp = torch.empty_like(wp)
        call("b2n_interlevel_fwd_bwd", ptr(c), ptr(w), ptr(cp), ptr(wp), R, Sc, Sp, 1.0, ptr(rows), ptr(d_wp), stream())
        ctx.save_for_backward(d_wp)
        return rows.sum()

    @staticmethod
    @_bwd
    def backward(ctx, g):
        (d_wp,) = ctx.saved_tensors
        return None, None, None, d_wp * g


def interlevel_loss(weights_list: List[Tensor], sdist_list: List[Tensor]) -> Tensor:
    """losses.py:113-132 — weights [R,S_i], sdist [R,S_i+1]; last entry is the (detached) target histogram."""
    c, w = sdist_list[-1].detach(), weights_list[-1].detach()
    total = 0.0
    for sd, wp in zip(sdist_list[:-1], weights_list[:-1]):
        total = total + _InterlevelFn.apply(c, w, sd.detach(), wp) / float(w.shape[0] * w.shape[1])
    return total


class _DistortionFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, t, w):
        t, w = _c(t), _c(w)
        R, S = w.shape
        rows = torch.empty(R, device=w.device, dtype=torch.float32)
        d_w = torch.empty_like(w)
        call("b2n_distortion_fwd_bwd", ptr(t), ptr(w), R, S, 1.0, ptr(rows), ptr(d_w), stream())
        ctx.save_for_backward(d_w)
        return rows.sum()

    @staticmethod
    @_bwd
    def backward(ctx, g):
        (d_w,) = ctx.saved_tensors
        return None, d_w * g


def distortion_loss(weights: Tensor, sdist: Tensor) -> Tensor:
    """losses.py:135-154 on the last level: mean over rays."""
    return _DistortionFn.apply(sdist.detach(), weights) / float(weights.shape[0])


# ----------------------------------------------------------------------------------------
# ray generation / colliders
# ----------------------------------------------------------------------------------------
def generate_rays(c2w: Tensor, intrinsics: Tensor, distortion: Optional[Tensor], ray_indices: Tensor):
    """c2w [C,3,4], intrinsics [C,4]=(fx,fy,cx,cy), distortion [C,6]|None, ray_indices int64 [R,3]."""
    c2w, intr, ri = _c(c2w.float()), _c(intrinsics.float()), _c(ray_indices.long())
    dist = _c(distortion.float()) if distortion is not None else None
    R = ri.shape[0]
    dev = c2w.device
    o = torch.empty(R, 3, device=dev)
    d = torch.empty(R, 3, device=dev)
    area = torch.empty(R, 1, device=dev)
    nrm = torch.empty(R, 1, device=dev)
    cam = torch.empty(R, 1, device=dev, dtype=torch.int64)
    call("b2n_raygen", ptr(c2w), ptr(intr), ptr(dist), ptr(ri, torch.int64), R, ptr(o), ptr(d), ptr(area), ptr(nrm),
         ptr(cam, torch.int64), stream())
    return dict(origins=o, directions=d, pixel_area=area, directions_norm=nrm, camera_indices=cam)


def generate_rays_coords(c2w: Tensor, intrinsics: Tensor, distortion: Optional[Tensor], cam_idx: Tensor,
                         coords: Optional[Tensor], height: int = 0, width: int = 0, cam_opt: Optional[Tensor] = None,
                         dist_delta: Optional[Tensor] = None):
    """Cameras.generate_rays on flattened arguments.  coords [R,2]=(y,x) with cam_idx [R]; or coords None for the whole
    images of cam_idx [K] -> R = height*width*K rays laid out [height, width, K]."""
    c2w, intr, ci = _c(c2w.float()), _c(intrinsics.float()), _c(cam_idx.long().reshape(-1))
    dist = _c(distortion.float()) if distortion is not None else None
    K = ci.shape[0]
    if coords is not None:
        coords = _c(coords.float().reshape(-1, 2))
        R = coords.shape[0]
        if K != R:
            raise ValueError("camera_indices and coords must have the same number of rays")
    else:
        R = K * height * width
    f = lambda t, w: None if t is None else _c(t.float().reshape(-1, w))
    opt, dd = f(cam_opt, 12), f(dist_delta, 6)
    dev = c2w.device
    o, d = torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev)
    area, nrm = torch.empty(R, 1, device=dev), torch.empty(R, 1, device=dev)
    cam = torch.empty(R, 1, device=dev, dtype=torch.int64)
    call("b2n_raygen_coords", ptr(c2w), ptr(intr), ptr(dist), ptr(ci, torch.int64), ptr(coords), R, K, int(height), int(width),
         ptr(opt), ptr(dd), ptr(o), ptr(d), ptr(area), ptr(nrm), ptr(cam, torch.int64), stream())
    return dict(origins=o, directions=d, pixel_area=area, directions_norm=nrm, camera_indices=cam)


def aabb_collide(origins: Tensor, directions: Tensor, aabb: Sequence[float], near_plane: float):
    o, d = _c(origins.float()), _c(directions.float())
    R = o.shape[0]
    n = torch.empty(R, 1, device=o.device)
    f = torch.empty(R, 1, device=o.device)
    box = host_floats(aabb)
    call("b2n_aabb_collide", ptr(o), ptr(d), C.cast(box, C.c_void_p), float(near_plane), R, ptr(n), ptr(f), stream())
    return n, f


class _PoseApplyFn(torch.autograd.Function):
    """CameraOptimizer.apply_to_raybundle (SO3xR3): gradients flow to pose_adjustment only (rays are data)."""

    @staticmethod
    @_fwd
    def forward(ctx, pose, cams, frozen, origins, directions):
        pose, o, d = _c(pose), _c(origins.float()), _c(directions.float())
        cams = _c(cams.reshape(-1).to(torch.int64))
        R = o.shape[0]
        out_o, out_d = torch.empty_like(o), torch.empty_like(d)
        call("b2n_pose_apply_fwd", ptr(pose), ptr(cams, torch.int64), ptr(frozen, torch.uint8), ptr(o), ptr(d), R,
             ptr(out_o), ptr(out_d), stream())
        ctx.save_for_backward(pose, cams, d)
        ctx.frozen = frozen
        return out_o, out_d

    @staticmethod
    @_bwd
    def backward(ctx, g_o, g_d):
        pose, cams, d = ctx.saved_tensors
        d_pose = torch.zeros_like(pose)
        g_o = None if g_o is None else _c(g_o.float())
        g_d = None if g_d is None else _c(g_d.float())
        if g_o is not None or g_d is not None:
            call("b2n_pose_apply_bwd", ptr(pose), ptr(cams, torch.int64), ptr(ctx.frozen, torch.uint8), ptr(d), ptr(g_o),
                 ptr(g_d), d.shape[0], ptr(d_pose), stream())
        return d_pose, None, None, None, None


def pose_apply(pose_adjustment: Tensor, camera_indices: Tensor, origins: Tensor, directions: Tensor,
               frozen: Optional[Tensor] = None):
    """(origins + t[cam], R(w[cam]) @ directions) for pose_adjustment [C,6] = (t | w); frozen: uint8 [C] or None."""
    return _PoseApplyFn.apply(pose_adjustment, camera_indices, frozen, origins, directions)


# ----------------------------------------------------------------------------------------
# packed path
# ----------------------------------------------------------------------------------------
def pack_info(ray_indices: Tensor, n_rays: int) -> Tensor:
    ri = _c(ray_indices.long())
    info = torch.empty(n_rays, 2, device=ri.device, dtype=torch.int64)
    call("b2n_pack_info", ptr(ri, torch.int64), ri.shape[0], n_rays, ptr(info, torch.int64), stream())
    return info


class _PackedWeightsFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, t_starts, t_ends, sigmas, packed_info):
        ts, te, sg = _c(t_starts), _c(t_ends), _c(sigmas)
        w, tr, al = torch.empty_like(sg), torch.empty_like(sg), torch.empty_like(sg)
        call("b2n_packed_weights_fwd", ptr(ts), ptr(te), ptr(sg), ptr(packed_info, torch.int64), packed_info.shape[0],
             ptr(w), ptr(tr), ptr(al), stream())
        ctx.save_for_backward(ts, te, sg, packed_info)
        ctx.mark_non_differentiable(tr, al)
        return w, tr, al

    @staticmethod
    @_bwd
    def backward(ctx, dw, _dt, _da):
        ts, te, sg, info = ctx.saved_tensors
        ds = torch.empty_like(sg)
        call("b2n_packed_weights_bwd", ptr(ts), ptr(te), ptr(sg), ptr(info, torch.int64), ptr(_c(dw.float())),
             info.shape[0], ptr(ds), stream())
        return None, None, ds, None


def packed_weights(t_starts, t_ends, sigmas, packed_info):
    return _PackedWeightsFn.apply(t_starts, t_ends, sigmas, _c(packed_info.long()))


class _PackedAccumFn(torch.autograd.Function):
    @staticmethod
    @_fwd
    def forward(ctx, weights, values, packed_info):
        w = _c(weights)
        v = _c(values) if values is not None else None
        d = 1 if v is None else v.shape[-1]
        R = packed_info.shape[0]
        out = torch.empty(R, d, device=w.device, dtype=torch.float32)
        call("b2n_packed_accumulate_fwd", ptr(w), ptr(v), d, ptr(packed_info, torch.int64), R, ptr(out), stream())
        ctx.d = d
        ctx.has_values = v is not None
        ctx.save_for_backward(w, v if v is not None else w, packed_info)
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout):
        w, v, info = ctx.saved_tensors
        v = v if ctx.has_values else None
        dw = torch.empty_like(w)
        dv = torch.empty_like(v) if v is not None else None
        call("b2n_packed_accumulate_bwd", ptr(w), ptr(v), ctx.d, ptr(info, torch.int64), ptr(_c(dout.float())),
             info.shape[0], ptr(dw), ptr(dv), stream())
        return dw, dv, None


def packed_accumulate(weights: Tensor, values: Optional[Tensor], packed_info: Tensor) -> Tensor:
    return _PackedAccumFn.apply(weights, values, _c(packed_info.long()))


def _u8(binaries: Tensor) -> Tensor:
    """bool and uint8 share the 1-byte 0/1 storage: view, never copy (the 128^3 x 4 grid is 8 MB)."""
    b = _c(binaries)
    return b.view(torch.uint8) if b.dtype == torch.bool else b


def scan_counts(counts: Tensor) -> Tuple[Tensor, int]:
    """int32 [n] -> (exclusive offsets int64 [n], total).  The total comes back to the host (one sync): the packed
    outputs of the reference's API are dynamically sized."""
    n = counts.shape[0]
    offsets = torch.empty(n, device=counts.device, dtype=torch.int64)
    total = torch.empty(1, device=counts.device, dtype=torch.int64)
    if n > 8192:
        blocks = (n + 1023) // 1024
        sums = torch.empty(blocks, device=counts.device, dtype=torch.int32)
        offs = torch.empty(blocks, device=counts.device, dtype=torch.int64)
        call("b2n_scan_counts_ws", ptr(counts, torch.int32), n, ptr(offsets, torch.int64), ptr(total, torch.int64),
             ptr(sums, torch.int32), ptr(offs, torch.int64), stream())
    else:
        call("b2n_scan_counts", ptr(counts, torch.int32), n, ptr(offsets, torch.int64), ptr(total, torch.int64), stream())
    return offsets, int(total.item())


def occgrid_march(origins, directions, binaries, roi_aabb, step, near_plane=0.0, far_plane=1e10, cone_angle=0.0,
                  jitter=None, t_min=None, t_max=None):
    """-> (ray_indices int64 [M], t_starts [M], t_ends [M]), sorted by ray then t.  binaries uint8/bool [levels,r,r,r]."""
    o, d = _c(origins.float()), _c(directions.float())
    R = o.shape[0]
    b = _u8(binaries)
    levels, res = b.shape[0], b.shape[1]
    roi = host_floats(roi_aabb)
    roip = C.cast(roi, C.c_void_p)
    f = lambda t: None if t is None else _c(t.float().reshape(-1))
    jit, tmn, tmx = f(jitter), f(t_min), f(t_max)
    counts = torch.empty(R, device=o.device, dtype=torch.int32)
    call("b2n_occgrid_count", ptr(o), ptr(d), ptr(tmn), ptr(tmx), ptr(b, torch.uint8), levels, res, roip, float(step),
         float(cone_angle), float(near_plane), float(far_plane), ptr(jit), R, ptr(counts, torch.int32), stream())
    offsets, M = scan_counts(counts) if R > 0 else (None, 0)
    ri = torch.empty(M, device=o.device, dtype=torch.int64)
    ts = torch.empty(M, device=o.device, dtype=torch.float32)
    te = torch.empty(M, device=o.device, dtype=torch.float32)
    if M > 0:
        call("b2n_occgrid_fill", ptr(o), ptr(d), ptr(tmn), ptr(tmx), ptr(b, torch.uint8), levels, res, roip, float(step),
             float(cone_angle), float(near_plane), float(far_plane), ptr(jit), R, ptr(offsets, torch.int64),
             ptr(ri, torch.int64), ptr(ts), ptr(te), stream())
    return ri, ts, te


def packed_positions(origins: Tensor, directions: Tensor, ray_indices: Tensor, t_starts: Tensor, t_ends: Tensor) -> Tensor:
    """Sample midpoints o[ray] + d[ray] * (ts + te)/2 -> [M,3] (no [M,3] gathers of origins/directions materialised)."""
    o, d, ri = _c(origins.float()), _c(directions.float()), _c(ray_indices.long())
    ts, te = _c(t_starts.float().reshape(-1)), _c(t_ends.float().reshape(-1))
    x = torch.empty(ri.shape[0], 3, device=o.device, dtype=torch.float32)
    call("b2n_packed_positions", ptr(o), ptr(d), ptr(ri, torch.int64), ptr(ts), ptr(te), ri.shape[0], ptr(x), stream())
    return x


def packed_prune(ray_indices: Tensor, t_starts: Tensor, t_ends: Tensor, trans: Tensor, alphas: Tensor, packed_info: Tensor,
                 early_stop_eps: float, alpha_thre: float, alpha_cap: Optional[Tensor] = None):
    """K8: keep samples with T >= early_stop_eps and alpha >= alpha_thre -> compacted (ray_indices, t_starts, t_ends).
    alpha_cap: optional device scalar; the threshold used is min(alpha_thre, alpha_cap) (nerfacc caps it with
    occs.mean() — kept on the device so no host round trip is needed)."""
    info = _c(packed_info.long())
    R = info.shape[0]
    tr, al, ts, te = _c(trans.float()), _c(alphas.float()), _c(t_starts.float()), _c(t_ends.float())
    counts = torch.empty(R, device=tr.device, dtype=torch.int32)
    call("b2n_packed_prune_count", ptr(tr), ptr(al), ptr(info, torch.int64), R, float(early_stop_eps), float(alpha_thre),
         ptr(alpha_cap), ptr(counts, torch.int32), stream())
    offsets, M = scan_counts(counts)
    ri = torch.empty(M, device=tr.device, dtype=torch.int64)
    ots, ote = torch.empty(M, device=tr.device), torch.empty(M, device=tr.device)
    if M > 0:
        call("b2n_packed_prune_fill", ptr(tr), ptr(al), ptr(info, torch.int64), R, float(early_stop_eps), float(alpha_thre),
             ptr(alpha_cap), ptr(offsets, torch.int64), ptr(ts), ptr(te), ptr(ri, torch.int64), ptr(ots), ptr(ote), stream())
    return ri, ots, ote


def occgrid_points(cell_ids: Optional[Tensor], jitter: Tensor, res: int, level_aabb: Sequence[float]) -> Tensor:
    """Jittered sample point of each listed cell of one grid level (cell_ids None = every cell in order)."""
    jit = _c(jitter.float())
    n = jit.shape[0]
    ids = None if cell_ids is None else _c(cell_ids.long())
    x = torch.empty(n, 3, device=jit.device, dtype=torch.float32)
    box = host_floats(level_aabb)
    call("b2n_occgrid_points", ptr(ids, torch.int64), ptr(jit), n, int(res), C.cast(box, C.c_void_p), ptr(x), stream())
    return x


def occgrid_ema(occs: Tensor, cell_ids: Optional[Tensor], occ_new: Tensor, level_offset: int, ema_decay: float) -> None:
    """In place: occs[level_offset + cell] = max(occs[..] * decay, occ_new) (old values; duplicates -> largest)."""
    occ_new = _c(occ_new.float().reshape(-1))
    n = occ_new.shape[0]
    ids = None if cell_ids is None else _c(cell_ids.long())
    scratch = torch.empty(n, device=occs.device, dtype=torch.float32)
    call("b2n_occgrid_ema", ptr(occs), ptr(ids, torch.int64), ptr(occ_new), n, int(level_offset), float(ema_decay),
         ptr(scratch), stream())


def occgrid_binarize(occs: Tensor, occ_thre: float, binaries_out: Optional[Tensor]) -> Tensor:
    """binaries_out (bool/uint8, occs.numel() elements; None = statistics only) <- occs > min(mean(occs), occ_thre).
    Returns a device tensor [threshold used, mean(occs)]."""
    parts = torch.empty(512, device=occs.device, dtype=torch.float64)
    stats = torch.empty(2, device=occs.device, dtype=torch.float32)
    b = None if binaries_out is None else _u8(binaries_out)
    call("b2n_occgrid_binarize", ptr(occs), occs.numel(), float(occ_thre), ptr(parts, torch.float64), ptr(stats),
         ptr(b, torch.uint8), stream())
    return stats


def ray_aabb_intersect(origins: Tensor, directions: Tensor, aabbs: Tensor, near_plane: float, far_plane: float,
                       miss_value: float):
    o, d, bx = _c(origins.float()), _c(directions.float()), _c(aabbs.float().reshape(-1, 6))
    n, k = o.shape[0], bx.shape[0]
    tmin = torch.empty(n, k, device=o.device, dtype=torch.float32)
    tmax = torch.empty_like(tmin)
    hits = torch.empty(n, k, device=o.device, dtype=torch.uint8)
    call("b2n_ray_aabb_intersect", ptr(o), ptr(d), ptr(bx), n, k, float(near_plane), float(far_plane), float(miss_value),
         ptr(tmin), ptr(tmax), ptr(hits, torch.uint8), stream())
    return tmin, tmax, hits.bool()


# ----------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-15,
              grad_scale: float = 1.0) -> None:
    call("b2n_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), int(step), float(lr), float(betas[0]),
         float(betas[1]), float(eps), float(grad_scale), stream())
