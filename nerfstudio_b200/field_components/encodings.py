"""Encodings behind the reference's `Encoding` interface, running sm_100a kernels.

Mirror of nerfstudio/field_components/encodings.py: `Encoding` (:36-55), `NeRFEncoding` (:90-186),
`HashEncoding` (:307-463), `SHEncoding` (:752-799).  Constructor arguments, public attributes, `get_out_dim`,
error behaviour (ValueError on bad arguments) and state_dict keys (`hash_table`) match the reference, so a
checkpoint written by either loads in the other.

`implementation="torch"` selects the torch-path arithmetic of the reference (the parity oracle, fp32 table of
`num_levels * 2**log2_hashmap_size` rows); `implementation="tcnn"` selects tiny-cuda-nn's grid semantics (dense
coarse levels, +0.5 offset) through `nerfstudio_b200.shims.tinycudann`.  Both run the same CUDA kernels; neither
falls back to PyTorch ops.
"""
from __future__ import annotations

from typing import Literal, Optional

import numpy as np
import torch
from torch import Tensor, nn

from .. import functional as F


class Encoding(nn.Module):
    def __init__(self, in_dim: int) -> None:
        if in_dim <= 0:
            raise ValueError("Input dimension should be greater than zero")
        super().__init__()
        self.in_dim = in_dim

    def get_out_dim(self) -> int:
        raise NotImplementedError

    def forward(self, in_tensor: Tensor) -> Tensor:
        raise NotImplementedError


class Identity(Encoding):
    def get_out_dim(self) -> int:
        return self.in_dim

    def forward(self, in_tensor: Tensor) -> Tensor:
        return in_tensor


class NeRFEncoding(Encoding):
    """sin/cos frequency encoding, out = D * F * 2 (+ D when include_input)."""

    def __init__(self, in_dim: int, num_frequencies: int, min_freq_exp: float, max_freq_exp: float,
                 include_input: bool = False, implementation: Literal["tcnn", "torch"] = "torch") -> None:
        super().__init__(in_dim)
        self.num_frequencies, self.min_freq, self.max_freq = num_frequencies, min_freq_exp, max_freq_exp
        self.include_input = include_input
        # the frequency table is evaluated once with torch, exactly as the reference does per call
        self._freqs = tuple((2 ** torch.linspace(min_freq_exp, max_freq_exp, num_frequencies)).tolist())

    def get_out_dim(self) -> int:
        return self.in_dim * self.num_frequencies * 2 + (self.in_dim if self.include_input else 0)

    def forward(self, in_tensor: Tensor, covs: Optional[Tensor] = None) -> Tensor:
        if covs is not None:
            raise NotImplementedError("integrated (mip-NeRF) encodings are outside the BASELINE hot path")
        flat = in_tensor.reshape(-1, self.in_dim)
        return F.freq_encode(flat, self._freqs, self.include_input).view(*in_tensor.shape[:-1], -1)


class HashEncoding(Encoding):
    """Multiresolution hash grid (Instant-NGP), gather/scatter on the B200 kernels."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19,
                 features_per_level: int = 2, hash_init_scale: float = 0.001,
                 implementation: Literal["tcnn", "torch"] = "tcnn",
                 interpolation: Optional[Literal["Nearest", "Linear", "Smoothstep"]] = None) -> None:
        super().__init__(in_dim=3)
        if interpolation not in (None, "Linear"):
            raise ValueError(f"interpolation '{interpolation}' is not supported (only Linear)")
        self.num_levels, self.min_res, self.features_per_level = num_levels, min_res, features_per_level
        self.hash_init_scale, self.log2_hashmap_size = hash_init_scale, log2_hashmap_size
        self.hash_table_size = 2 ** log2_hashmap_size
        levels = torch.arange(num_levels)
        self.growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
        # float32 pow + floor, bit-identical to the reference's table (last nerfacto level is 2047, not 2048)
        self.scalings = torch.floor(min_res * self.growth_factor ** levels)
        self.hash_offset = levels * self.hash_table_size
        self.implementation = implementation
        self.tcnn_encoding = None
        self.hash_table = torch.empty(0)
        if implementation == "torch":
            self.grid = F.GridSpec(self.scalings.tolist(), log2_hashmap_size, features_per_level, "torch")
            table = (torch.rand(self.hash_table_size * num_levels, features_per_level) * 2 - 1) * hash_init_scale
            self.hash_table = nn.Parameter(table)
        elif implementation == "tcnn":
            from ..shims import tinycudann as tcnn

            self.tcnn_encoding = tcnn.Encoding(
                n_input_dims=3,
                encoding_config=self.get_tcnn_encoding_config(num_levels, features_per_level, log2_hashmap_size,
                                                              min_res, self.growth_factor, interpolation))
        else:
            raise ValueError(f"unknown implementation {implementation!r}")

    @classmethod
    def get_tcnn_encoding_config(cls, num_levels, features_per_level, log2_hashmap_size, min_res, growth_factor,
                                 interpolation=None) -> dict:
        cfg = {"otype": "HashGrid", "n_levels": num_levels, "n_features_per_level": features_per_level,
               "log2_hashmap_size": log2_hashmap_size, "base_resolution": min_res, "per_level_scale": growth_factor}
        if interpolation is not None:
            cfg["interpolation"] = interpolation
        return cfg

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def forward(self, in_tensor: Tensor) -> Tensor:
        if self.tcnn_encoding is not None:
            return self.tcnn_encoding(in_tensor)
        flat = in_tensor.reshape(-1, 3)
        return F.hash_encode(flat, self.hash_table, self.grid).view(*in_tensor.shape[:-1], self.get_out_dim())


class SHEncoding(Encoding):
    """Real spherical harmonics of the (already [0,1]-mapped) directions; levels = degree + 1."""

    def __init__(self, levels: int = 4, implementation: Literal["tcnn", "torch"] = "torch") -> None:
        super().__init__(in_dim=3)
        if levels <= 0 or levels > 5:
            raise ValueError(f"Spherical harmonic encoding only supports 1 to 5 levels, requested {levels}")
        self.levels = levels
        self.implementation = implementation
        self.tcnn_encoding = None
        if implementation == "tcnn":
            from ..shims import tinycudann as tcnn

            self.tcnn_encoding = tcnn.Encoding(n_input_dims=3, encoding_config=self.get_tcnn_encoding_config(levels))

    @classmethod
    def get_tcnn_encoding_config(cls, levels: int) -> dict:
        return {"otype": "SphericalHarmonics", "degree": levels}

    def get_out_dim(self) -> int:
        return self.levels ** 2

    def forward(self, in_tensor: Tensor) -> Tensor:
        if self.tcnn_encoding is not None:
            return self.tcnn_encoding(in_tensor)
        flat = in_tensor.reshape(-1, 3)
        return F.sh_encode(flat, self.levels, remap01=False).view(*in_tensor.shape[:-1], -1)
