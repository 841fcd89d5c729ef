"""ctypes binding of libb200nerf.so (include/b200nerf.h).  No torch types cross this boundary: tensors are
passed as raw device pointers + sizes and the current CUDA stream as a void*.

The product path fails loudly: if the shared library is missing, cannot be loaded, or a call returns an
error code, a RuntimeError/ValueError is raised.  There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200nerf.so")

MAX_LEVELS = 32
MAX_MLP_LAYERS = 10

GRID_TORCH, GRID_TCNN = 0, 1
ACT = {"none": 0, "relu": 1, "sigmoid": 2, "softplus": 3, "tanh": 4}
SPACING = {"uniform": 0, "piecewise": 1, "lindisp": 2, "sqrt": 3, "log": 4}
BG_NONE, BG_LAST_SAMPLE, BG_CONSTANT = 0, 1, 2


class B2nGrid(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int32), ("n_features", C.c_int32), ("log2_hashmap_size", C.c_int32), ("mode", C.c_int32),
        ("scale", C.c_float * MAX_LEVELS), ("resolution", C.c_uint32 * MAX_LEVELS),
        ("offset", C.c_uint32 * MAX_LEVELS), ("size", C.c_uint32 * MAX_LEVELS), ("hashed", C.c_uint32 * MAX_LEVELS),
    ]


class B2nMlp(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("in_dim", C.c_int32), ("hidden_act", C.c_int32), ("out_act", C.c_int32),
        ("out_dims", C.c_int32 * MAX_MLP_LAYERS), ("skip", C.c_int32 * MAX_MLP_LAYERS),
        ("w", C.c_void_p * MAX_MLP_LAYERS), ("b", C.c_void_p * MAX_MLP_LAYERS),
    ]


class B2nMlpGrad(C.Structure):
    _fields_ = [("dw", C.c_void_p * MAX_MLP_LAYERS), ("db", C.c_void_p * MAX_MLP_LAYERS)]


_P, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# name -> argtypes (return type is always int unless listed in _RET)
_SIGNATURES = {
    "b2n_device_info": [_P],
    "b2n_tune": [C.c_char_p, C.c_int],
    "b2n_hashgrid_fwd": [C.POINTER(B2nGrid), _P, _P, _I64, _P, _P, _P],
    "b2n_hashgrid_bwd": [C.POINTER(B2nGrid), _P, _P, _P, _I64, _P, _P, _P],
    "b2n_hashgrid_dx": [C.POINTER(B2nGrid), _P, _P, _P, _I64, _P, _P],
    "b2n_mlp_fwd": [C.POINTER(B2nMlp), _P, _I64, _P, _P, _P],
    "b2n_mlp_bwd": [C.POINTER(B2nMlp), C.POINTER(B2nMlpGrad), _P, _P, _P, _P, _I64, _P, _P],
    "b2n_mlp_tc_fwd": [C.POINTER(B2nMlp), _P, _I64, _I64, _P, _P, _P],
    "b2n_mlp_tc_bwd": [C.POINTER(B2nMlp), C.POINTER(B2nMlpGrad), _P, _I64, _P, _P, _P, _I64, _P, _I64, _P],
    "b2n_mlp_tc_pack": [C.POINTER(B2nMlp), _P, _P],
    "b2n_mlp_tc_fwd_ws": [C.POINTER(B2nMlp), _P, _I64, _I64, _P, _P, _P, _P],
    "b2n_mlp_tc_bwd_ws": [C.POINTER(B2nMlp), C.POINTER(B2nMlpGrad), _P, _I64, _P, _P, _P, _I64, _P, _I64, _P, _P],
    "b2n_linear_fwd": [_P, _I64, _I32, _I64, _P, _P, _I32, _I32, _P, _P],
    "b2n_linear_bwd": [_P, _I64, _I32, _I64, _P, _I32, _I32, _P, _P, _P, _P, _I64, _P, _P, _P],
    "b2n_sh_fwd": [_P, _I64, _I32, _I32, _P, _P],
    "b2n_freq_fwd": [_P, _I64, _I32, _P, _I32, _I32, _P, _P],
    "b2n_freq_bwd": [_P, _P, _I64, _I32, _P, _I32, _I32, _P, _P],
    "b2n_positions_fwd": [_P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P],
    "b2n_positions_bwd": [_P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _P, _I32, _P, _P, _P],
    "b2n_density_act_fwd": [_P, _I64, _P, _I64, _F, _P, _P],
    "b2n_density_act_bwd": [_P, _I64, _P, _P, _I64, _F, _P, _I64, _P],
    "b2n_spaced_sample": [_P, _P, _P, _P, _I32, _I64, _I32, _I32, _P, _P, _P],
    "b2n_pdf_sample": [_P, _P, _P, _P, _I32, _P, _P, _I64, _I32, _I32, _F, _P, _F, _F, _I32, _P, _P, _P, _P, _P],
    "b2n_torch_row_sum": [_P, _I64, _I32, _P, _P],
    "b2n_weights_fwd": [_P, _P, _I64, _P, _I64, _I32, _P, _P],
    "b2n_weights_bwd": [_P, _P, _I64, _P, _P, _I64, _I32, _P, _P],
    "b2n_composite_fwd": [_P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P],
    "b2n_composite_bwd": [_P, _P, _P, _P, _I64, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P],
    "b2n_interlevel_fwd_bwd": [_P, _P, _P, _P, _I64, _I32, _I32, _F, _P, _P, _P],
    "b2n_distortion_fwd_bwd": [_P, _P, _I64, _I32, _F, _P, _P, _P],
    "b2n_raygen": [_P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P],
    "b2n_raygen_coords": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P],
    "b2n_pixel_sample_raygen": [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P],
    "b2n_aabb_collide": [_P, _P, _P, _F, _I64, _P, _P, _P],
    "b2n_pose_apply_fwd": [_P, _P, _P, _P, _P, _I64, _P, _P, _P],
    "b2n_pose_apply_bwd": [_P, _P, _P, _P, _P, _P, _I64, _P, _P],
    "b2n_pack_info": [_P, _I64, _I64, _P, _P],
    "b2n_packed_weights_fwd": [_P, _P, _P, _P, _I64, _P, _P, _P, _P],
    "b2n_packed_weights_bwd": [_P, _P, _P, _P, _P, _I64, _P, _P],
    "b2n_packed_accumulate_fwd": [_P, _P, _I32, _P, _I64, _P, _P],
    "b2n_packed_accumulate_bwd": [_P, _P, _I32, _P, _P, _I64, _P, _P, _P],
    "b2n_occgrid_count": [_P, _P, _P, _P, _P, _I32, _I32, _P, _F, _F, _F, _F, _P, _I64, _P, _P],
    "b2n_occgrid_fill": [_P, _P, _P, _P, _P, _I32, _I32, _P, _F, _F, _F, _F, _P, _I64, _P, _P, _P, _P, _P],
    "b2n_scan_counts": [_P, _I64, _P, _P, _P],
    "b2n_scan_counts_ws": [_P, _I64, _P, _P, _P, _P, _P],
    "b2n_packed_positions": [_P, _P, _P, _P, _P, _I64, _P, _P],
    "b2n_packed_prune_count": [_P, _P, _P, _I64, _F, _F, _P, _P, _P],
    "b2n_packed_prune_fill": [_P, _P, _P, _I64, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P],
    "b2n_occgrid_points": [_P, _P, _I64, _I32, _P, _P, _P],
    "b2n_occgrid_ema": [_P, _P, _P, _I64, _I64, _F, _P, _P],
    "b2n_occgrid_binarize": [_P, _I64, _F, _P, _P, _P, _P],
    "b2n_ray_aabb_intersect": [_P, _P, _P, _I64, _I32, _F, _F, _F, _P, _P, _P, _P],
    "b2n_density_field_fwd": [C.POINTER(B2nGrid), C.POINTER(B2nMlp), _P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _F, _P, _P],
    "b2n_density_field_bwd": [C.POINTER(B2nGrid), C.POINTER(B2nMlp), C.POINTER(B2nMlpGrad), _P, _P, _P, _P, _P, _I64, _I64,
                              _I32, _I32, _P, _F, _P, _P, _P],
    "b2n_density_field_bwd_ws": [C.POINTER(B2nGrid), C.POINTER(B2nMlp), C.POINTER(B2nMlpGrad), _P, _P, _P, _P, _P, _I64, _I64,
                                 _I32, _I32, _P, _F, _P, _P, _P, _P],
    "b2n_density_field_bwd_rays": [C.POINTER(B2nGrid), C.POINTER(B2nMlp), C.POINTER(B2nMlpGrad), _P, _P, _P, _P, _P, _I64,
                                   _I64, _I32, _I32, _P, _F, _P, _P, _P, _P, _P, _P],
    "b2n_pose_regularizer": [_P, _I32, _F, _F, _F, _P, _P, _P],
    "b2n_gs_project_fwd": [_P, _P, _P, _P, _I32, _I32, _I64, _P, _P, _I32, _I32, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "b2n_gs_project_bwd": [_P, _P, _P, _P, _I32, _I32, _I64, _P, _P, _I32, _I32, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P,
                           _P, _P, _P, _P, _P],
    "b2n_gs_emit": [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P],
    "b2n_gs_tile_ranges": [_P, _I64, _P, _P, _P],
    "b2n_gs_rasterize_fwd": [_I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "b2n_gs_rasterize_bwd": [_I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "b2n_tc_selftest": [_I32, _I32, _P, _I32, _I32, _P, _I32, _I32, _I32, _I32, _I32, _P, _P],
    "b2n_tc_timing": [_I32, _I32, _I32, _P, _P],
    "b2n_adam_step_dev": [_P, _P, _P, _P, _I64, _P, C.c_double, C.c_double, C.c_double, _P],
    "b2n_head_input_fwd": [_P, _I32, _P, _I32, _I32, _P, _P, _I32, _I32, _I64, _I32, _P, _I32, _P],
    "b2n_head_input_bwd": [_P, _I32, _I32, _I32, _I32, _P, _P, _I64, _I32, _P, _I32, _P, _P],
    "b2n_head_input_fwd_part": [_P, _I32, _P, _I32, _I32, _P, _P, _I32, _I32, _I64, _I32, _P, _I32, _I32, _P],
    "b2n_head_input_bwd_part": [_P, _I32, _I32, _I32, _I32, _P, _P, _I64, _I32, _P, _I32, _P, _I32, _P],
    "b2n_mse_fwd_bwd": [_P, _P, _I64, _F, _P, _P, _P],
    "b2n_sum_rows": [_P, _I64, _F, _P, _P],
    "b2n_nerfacto_ray_tail": [_I64, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _F, _P, _P, _I32, _P, _F, _F, _F, _P, _P, _P, _P, _P, _P,
                              _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "b2n_loss_finalize": [_P, _I64, _F, _F, _F, _P, _P],
    "b2n_weights_pdf_sample": [_P, _P, _P, _P, _P, _I32, _P, _P, _I64, _I32, _I32, _F, _P, _F, _F, _I32, _P, _P, _P, _P],
    "b2n_step_begin": [_P, _I64, _P, _P, _I64, _P, _I64, _P],
    "b2n_add_inplace": [_P, _P, _I64, _P],
    "b2n_loss_total": [_P, _I32, _P, _P, _P],
    "b2n_zero_async": [_P, _I64, _P],
    "b2n_adam_step": [_P, _P, _P, _P, _I64, _I32, C.c_double, C.c_double, C.c_double, C.c_double, _F, _P],
}
_RET = {"b2n_version": C.c_char_p, "b2n_last_error": C.c_char_p}
_RET_ARGS = {"b2n_mlp_tc_workspace_bytes": ([C.POINTER(B2nMlp)], C.c_int64)}

EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + list(_RET) + list(_RET_ARGS))

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library (never builds, never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m nerfstudio_b200.build` (nvcc, sm_100a). "
            "nerfstudio_b200 has no CPU/PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argtypes, C.c_int
    for name, ret in _RET.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = [], ret
    for name, (argtypes, ret) in _RET_ARGS.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argtypes, ret
    _lib = lib
    # B2N_TUNE="key=value,key=value": launch-geometry / kernel-variant knobs applied once at load (see b2n_tune)
    for kv in filter(None, os.environ.get("B2N_TUNE", "").split(",")):
        k, v = kv.split("=")
        if not lib.b2n_tune(k.encode(), int(v)):
            raise RuntimeError(f"B2N_TUNE: unknown tuning key {k!r}")
    return lib


def last_error() -> str:
    return load().b2n_last_error().decode()


def check(code: int, what: str) -> None:
    if code == 0:
        return
    msg = last_error()
    if code < 0:
        raise ValueError(f"{what}: {msg} (code {code})")
    raise RuntimeError(f"{what}: CUDA error {code}: {msg}")


LAUNCHES = 0  # kernel-launching C-ABI calls made by this process (bench.py reports it as gpu_launches)


PROFILE = None  # set to {} to time every C-ABI launch with CUDA events on the launching stream (eager mode only)
PROFILE_BY_SIZE = True  # False: one row per entry point (dynamic sizes, e.g. packed instant-ngp samples); n is summed
_N_ARG = {"b2n_hashgrid_fwd": 3, "b2n_hashgrid_bwd": 4, "b2n_hashgrid_dx": 4, "b2n_mlp_fwd": 2, "b2n_mlp_bwd": 6, "b2n_mlp_tc_fwd": 3,
          "b2n_mlp_tc_bwd": 7, "b2n_mlp_tc_fwd_ws": 3, "b2n_mlp_tc_bwd_ws": 7}


def call(name: str, *args) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if PROFILE is None:
        check(getattr(load(), name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(load(), name)(*args), name)
    e1.record()
    n_arg = int(args[_N_ARG[name]]) if name in _N_ARG else 0
    if not PROFILE_BY_SIZE:
        key = name
    elif name == "b2n_density_field_fwd":
        key = f"{name}[n={args[8] * args[9]}]"
    elif name in ("b2n_density_field_bwd", "b2n_density_field_bwd_ws", "b2n_density_field_bwd_rays"):
        key = f"b2n_density_field_bwd[n={args[9] * args[10]}]"
    elif name in _N_ARG and name.startswith("b2n_mlp"):
        m = args[0]._obj  # byref(B2nMlp): two networks of one model may share n (base / colour head)
        key = f"{name.replace('_ws', '')}[n={args[_N_ARG[name]]},in={m.in_dim},out={m.out_dims[m.n_layers - 1]}]"
    else:
        key = f"{name}[n={args[_N_ARG[name]]}]" if name in _N_ARG else name
    PROFILE.setdefault(key, []).append((e0, e1, n_arg))


def profile_summary() -> dict:
    """key -> (launches, total ms); call after torch.cuda.synchronize()."""
    return {k: (len(v), sum(a.elapsed_time(b) for a, b, _ in v)) for k, v in (PROFILE or {}).items()}


def profile_sizes() -> dict:
    """key -> sum of the size argument n over the profiled launches (entry points listed in _N_ARG)."""
    return {k: sum(n for _, _, n in v) for k, v in (PROFILE or {}).items()}


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = torch.float32) -> C.c_void_p:
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("nerfstudio_b200 kernels need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def host_floats(values) -> C.Array:
    vals = [float(v) for v in values]
    return (C.c_float * len(vals))(*vals)


def device_info() -> dict:
    out = (C.c_int32 * 4)()
    call("b2n_device_info", C.cast(out, C.c_void_p))
    return dict(sm_count=out[0], cc_major=out[1], cc_minor=out[2], max_smem_optin=out[3])


def tune(key: str, value: int) -> bool:
    return bool(load().b2n_tune(key.encode(), int(value)))
