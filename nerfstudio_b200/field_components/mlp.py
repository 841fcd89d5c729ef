"""MLP / MLPWithHashEncoding behind the reference's interface, on the fused sm_100a MLP kernel.

Mirror of nerfstudio/field_components/mlp.py:74-295: same constructor arguments, `layers` ModuleList of
nn.Linear (so state_dict keys are `layers.i.weight/bias`, reference default init), skip connections, hidden and
output activations named by nn.Module instances.  Networks that fit the fused kernel's shared-memory budget run
as ONE launch (all nerfacto / instant-ngp networks do); wider ones (vanilla-nerf's 8x256) run layer by layer
on library GEMMs — config 1 of BASELINE.json is the reference's CPU-runnable case, not a bench line.
"""
from __future__ import annotations

from typing import Literal, Optional, Set, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from .. import functional as F
from .encodings import HashEncoding


def _act_name(act: Optional[nn.Module]) -> str:
    if act is None:
        return "none"
    for cls, name in ((nn.ReLU, "relu"), (nn.Sigmoid, "sigmoid"), (nn.Softplus, "softplus"), (nn.Tanh, "tanh")):
        if isinstance(act, cls):
            return name
    raise ValueError(f"activation {act} is not supported by the fused MLP kernel")


class MLP(nn.Module):
    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: Optional[int] = None,
                 skip_connections: Optional[Tuple[int]] = None, activation: Optional[nn.Module] = nn.ReLU(),
                 out_activation: Optional[nn.Module] = None,
                 implementation: Literal["tcnn", "torch"] = "torch") -> None:
        super().__init__()
        assert in_dim > 0
        self.in_dim = in_dim
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers, self.layer_width = num_layers, layer_width
        self.skip_connections = skip_connections
        self._skip_connections: Set[int] = set(skip_connections) if skip_connections else set()
        self.activation, self.out_activation = activation, out_activation
        self.implementation = implementation
        self.tcnn_encoding = None
        if implementation == "tcnn":
            from ..shims import tinycudann as tcnn

            self.tcnn_encoding = tcnn.Network(
                n_input_dims=in_dim, n_output_dims=self.out_dim,
                network_config=self.get_tcnn_network_config(activation, out_activation, layer_width, num_layers))
            return
        dims = [layer_width] * (num_layers - 1) + [self.out_dim]
        layers, prev = [], in_dim
        for i, out in enumerate(dims):
            assert not (i == 0 and i in self._skip_connections), "Skip connection at layer 0 doesn't make sense."
            layers.append(nn.Linear(prev + (in_dim if i in self._skip_connections else 0), out))
            prev = out
        self.layers = nn.ModuleList(layers)
        self.spec = F.MlpSpec(in_dim, dims, skip=self._skip_connections, hidden_act=_act_name(activation),
                              out_act=_act_name(out_activation))
        self._fused = self.spec.fits_fused_kernel()

    @classmethod
    def get_tcnn_network_config(cls, activation, out_activation, layer_width, num_layers) -> dict:
        names = {"none": "None", "relu": "ReLU", "sigmoid": "Sigmoid", "softplus": "Softplus", "tanh": "Tanh"}
        return {"otype": "FullyFusedMLP" if layer_width in (16, 32, 64, 128) else "CutlassMLP",
                "activation": names[_act_name(activation)], "output_activation": names[_act_name(out_activation)],
                "n_neurons": layer_width, "n_hidden_layers": num_layers - 1}

    def get_out_dim(self) -> int:
        return self.out_dim

    def forward(self, in_tensor: Tensor) -> Tensor:
        if self.tcnn_encoding is not None:
            return self.tcnn_encoding(in_tensor)
        flat = in_tensor.reshape(-1, self.in_dim)
        if self._fused:
            out = F.mlp(self.spec, flat, [l.weight for l in self.layers], [l.bias for l in self.layers])
        else:
            out = self._wide_forward(flat)
        return out.view(*in_tensor.shape[:-1], self.out_dim)

    def _wide_forward(self, x: Tensor) -> Tensor:
        """Widths beyond the fused kernels' shared memory (vanilla-nerf 8 x 256 with a skip): one tiled fp32 GEMM launch per
        layer with the bias and activation fused into its epilogue (csrc/wide_mlp.cu) — no library GEMM."""
        h = x.float()
        last = len(self.layers) - 1
        for i, layer in enumerate(self.layers):
            if i in self._skip_connections:
                h = torch.cat([x.float(), h], -1)
            act = _act_name(self.out_activation) if i == last else _act_name(self.activation)
            h = F.linear(h, layer.weight, layer.bias, act)
        return h


class MLPWithHashEncoding(nn.Module):
    """Hash grid + tiny MLP.  `model` is `Sequential(HashEncoding, MLP)` (state_dict keys `model.0.hash_table`,
    `model.1.layers.i.*`) for implementation="torch", one tcnn-style module for "tcnn" — as in the reference."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19,
                 features_per_level: int = 2, hash_init_scale: float = 0.001,
                 interpolation: Optional[Literal["Nearest", "Linear", "Smoothstep"]] = None, num_layers: int = 2,
                 layer_width: int = 64, out_dim: Optional[int] = None, skip_connections: Optional[Tuple[int]] = None,
                 activation: Optional[nn.Module] = nn.ReLU(), out_activation: Optional[nn.Module] = None,
                 implementation: Literal["tcnn", "torch"] = "torch") -> None:
        super().__init__()
        self.in_dim = 3
        self.num_levels, self.min_res, self.max_res = num_levels, min_res, max_res
        self.features_per_level, self.hash_init_scale = features_per_level, hash_init_scale
        self.log2_hashmap_size, self.hash_table_size = log2_hashmap_size, 2 ** log2_hashmap_size
        self.growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers, self.layer_width = num_layers, layer_width
        self.skip_connections, self.activation, self.out_activation = skip_connections, activation, out_activation
        self.tcnn_encoding = None
        if implementation == "tcnn":
            from ..shims import tinycudann as tcnn

            self.model = tcnn.NetworkWithInputEncoding(
                n_input_dims=3, n_output_dims=self.out_dim,
                encoding_config=HashEncoding.get_tcnn_encoding_config(num_levels, features_per_level,
                                                                      log2_hashmap_size, min_res, self.growth_factor,
                                                                      interpolation),
                network_config=MLP.get_tcnn_network_config(activation, out_activation, layer_width, num_layers))
        else:
            encoder = HashEncoding(num_levels=num_levels, min_res=min_res, max_res=max_res,
                                   log2_hashmap_size=log2_hashmap_size, features_per_level=features_per_level,
                                   hash_init_scale=hash_init_scale, implementation="torch")
            mlp = MLP(in_dim=encoder.get_out_dim(), num_layers=num_layers, layer_width=layer_width, out_dim=out_dim,
                      skip_connections=skip_connections, activation=activation, out_activation=out_activation,
                      implementation="torch")
            self.model = torch.nn.Sequential(encoder, mlp)

    def get_out_dim(self) -> int:
        return self.out_dim

    def forward(self, in_tensor: Tensor) -> Tensor:
        return self.model(in_tensor)
