"""`tinycudann`-compatible module surface on the B200 kernels.

nerfstudio's `implementation="tcnn"` code paths construct `tcnn.Encoding`, `tcnn.Network` and
`tcnn.NetworkWithInputEncoding` (nerfstudio/field_components/encodings.py:129,362,774; mlp.py:110,252) and gate
them on `import tinycudann` succeeding (utils/external.py:38-58).  This module exports those three classes with
the same constructor/`forward` contract (one flat fp32 `params` Parameter, `n_output_dims`, JSON-style config
dicts), implemented with libb200nerf.so — so `TCNN_EXISTS` can be true on a B200 without tiny-cuda-nn.

Semantics follow tiny-cuda-nn's published behaviour: HashGrid with `scale = 2^(l*log2 g)*base-1`,
`+0.5` offset, dense coarse levels; bias-free MLPs; SphericalHarmonics on `2x-1` with tcnn's sign convention.
tiny-cuda-nn's sources are not available in this environment: numerical parity with it is UNPINNED (the oracle
for this mode is oracle/nerf_oracle.py:tcnn_hash_encode).  Unlike tcnn we keep fp32 tables and outputs (the
reference casts tcnn's fp16 outputs to fp32 immediately: nerfacto_field.py:230).

Parameter LAYOUT is not tiny-cuda-nn's either: the flat `params` holds the exact (out, in) weight matrices back to back,
without tcnn's padding of input/output widths to multiples of 16, so `params.numel()` differs and checkpoints written by
genuine tiny-cuda-nn modules do not load into these (nor the other way round): `load_state_dict` reports the size
mismatch.  Checkpoints of the torch-mode modules (`implementation="torch"`) are interchangeable with the reference's.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import Tensor, nn

from .. import functional as F

_ACT = {"None": "none", "ReLU": "relu", "Sigmoid": "sigmoid", "Softplus": "softplus", "Tanh": "tanh"}
_SH_SIGN = [1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1]


class _GridEncoding:
    def __init__(self, cfg: Dict):
        self.grid = F.GridSpec.tcnn(int(cfg["n_levels"]), int(cfg["base_resolution"]), float(cfg["per_level_scale"]),
                                    int(cfg["log2_hashmap_size"]), int(cfg["n_features_per_level"]))
        if cfg.get("interpolation", "Linear") != "Linear":
            raise NotImplementedError("only Linear interpolation is implemented")
        self.n_params = self.grid.n_rows * self.grid.n_features
        self.n_output_dims = self.grid.out_dim

    def init(self, params: Tensor) -> None:
        params.uniform_(-1e-4, 1e-4)

    def __call__(self, x: Tensor, params: Tensor) -> Tensor:
        return F.hash_encode(x, params.view(self.grid.n_rows, self.grid.n_features), self.grid)


class _SHEncoding:
    def __init__(self, cfg: Dict):
        self.levels = int(cfg["degree"])
        self.n_params, self.n_output_dims = 0, self.levels ** 2
        self.sign = None

    def init(self, params: Tensor) -> None:
        pass

    def __call__(self, x: Tensor, params: Tensor) -> Tensor:
        out = F.sh_encode(x * 2.0 - 1.0, self.levels, remap01=False)
        if self.sign is None or self.sign.device != out.device:
            self.sign = torch.tensor(_SH_SIGN[: self.levels ** 2], device=out.device, dtype=torch.float32)
        return out * self.sign


def _make_encoding(n_input_dims: int, cfg: Dict):
    otype = cfg["otype"]
    if otype in ("HashGrid", "Grid"):
        if n_input_dims != 3:
            raise NotImplementedError("grid encodings are implemented for 3-D inputs")
        return _GridEncoding(cfg)
    if otype == "SphericalHarmonics":
        return _SHEncoding(cfg)
    raise NotImplementedError(f"tcnn encoding otype {otype!r} is not implemented (outside the BASELINE hot path)")


class _Net:
    def __init__(self, n_in: int, n_out: int, cfg: Dict):
        hidden = int(cfg["n_hidden_layers"])
        width = int(cfg["n_neurons"])
        dims = [width] * hidden + [n_out]
        self.spec = F.MlpSpec(n_in, dims, hidden_act=_ACT[cfg.get("activation", "ReLU")],
                              out_act=_ACT[cfg.get("output_activation", "None")], bias=False)
        self.shapes = [(d, self.spec.layer_in(i)) for i, d in enumerate(dims)]
        self.n_params = sum(a * b for a, b in self.shapes)
        if not self.spec.fits_fused_kernel():
            raise NotImplementedError("network too wide for the fused kernel")

    def init(self, params: Tensor) -> None:
        off = 0
        for out, inn in self.shapes:
            bound = math.sqrt(6.0 / (inn + out))  # xavier uniform, as tcnn initialises its matrices
            params[off: off + out * inn].uniform_(-bound, bound)
            off += out * inn

    def __call__(self, x: Tensor, params: Tensor) -> Tensor:
        ws, off = [], 0
        for out, inn in self.shapes:
            ws.append(params[off: off + out * inn].view(out, inn))
            off += out * inn
        return F.mlp(self.spec, x, ws, [None] * len(ws))


class _Module(nn.Module):
    dtype = torch.float32
    loss_scale = 1.0

    def _finish(self, n_params: int, seed: int, inits) -> None:
        p = torch.zeros(max(n_params, 0))
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        off = 0
        for part, n in inits:
            part.init(p[off: off + n])
            off += n
        torch.random.set_rng_state(gen_state)
        self.params = nn.Parameter(p)

    @staticmethod
    def _flat(x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("tinycudann (B200 shim) modules only run on CUDA tensors")
        return x.reshape(-1, x.shape[-1]).float()


class Encoding(_Module):
    def __init__(self, n_input_dims: int, encoding_config: Dict, seed: int = 1337, dtype=None):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.enc = _make_encoding(n_input_dims, encoding_config)
        self.n_output_dims = self.enc.n_output_dims
        self._finish(self.enc.n_params, seed, [(self.enc, self.enc.n_params)])

    def forward(self, x: Tensor) -> Tensor:
        return self.enc(self._flat(x), self.params).view(*x.shape[:-1], self.n_output_dims)


class Network(_Module):
    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: Dict, seed: int = 1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.net = _Net(n_input_dims, n_output_dims, network_config)
        self._finish(self.net.n_params, seed, [(self.net, self.net.n_params)])

    def forward(self, x: Tensor) -> Tensor:
        return self.net(self._flat(x), self.params).view(*x.shape[:-1], self.n_output_dims)


class NetworkWithInputEncoding(_Module):
    """params = [network weights | grid features] (tcnn's order)."""

    def __init__(self, n_input_dims: int, n_output_dims: int, encoding_config: Dict, network_config: Dict,
                 seed: int = 1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.enc = _make_encoding(n_input_dims, encoding_config)
        self.net = _Net(self.enc.n_output_dims, n_output_dims, network_config)
        self._finish(self.net.n_params + self.enc.n_params, seed,
                     [(self.net, self.net.n_params), (self.enc, self.enc.n_params)])

    def forward(self, x: Tensor) -> Tensor:
        n = self.net.n_params
        h = self.enc(self._flat(x), self.params[n:])
        return self.net(h, self.params[:n]).view(*x.shape[:-1], self.n_output_dims)
