"""Field-head names and the two heads vanilla-nerf uses (mirror of nerfstudio/field_components/field_heads.py:28-119).

When nerfstudio itself is importable its own `FieldHeadNames` enum is re-exported, so dictionaries produced by
our fields are keyed exactly like the reference's (drop-in inside unmodified nerfstudio models)."""
from enum import Enum
from typing import Optional

from torch import Tensor, nn

try:  # pragma: no cover - exercised only where nerfstudio is installed
    from nerfstudio.field_components.field_heads import FieldHeadNames  # type: ignore
except Exception:  # noqa: BLE001

    class FieldHeadNames(Enum):
        RGB = "rgb"
        SH = "sh"
        DENSITY = "density"
        NORMALS = "normals"
        PRED_NORMALS = "pred_normals"
        UNCERTAINTY = "uncertainty"
        BACKGROUND_RGB = "background_rgb"
        TRANSIENT_RGB = "transient_rgb"
        TRANSIENT_DENSITY = "transient_density"
        SEMANTICS = "semantics"
        SDF = "sdf"
        ALPHA = "alpha"
        GRADIENT = "gradient"


class FieldHead(nn.Module):
    """Linear layer + activation producing one named field output."""

    def __init__(self, out_dim: int, field_head_name, in_dim: Optional[int] = None,
                 activation: Optional[nn.Module] = None) -> None:
        super().__init__()
        self.out_dim, self.activation, self.field_head_name = out_dim, activation, field_head_name
        self.net = None
        self.in_dim = None
        if in_dim is not None:
            self.set_in_dim(in_dim)

    def set_in_dim(self, in_dim: int) -> None:
        self.in_dim = in_dim
        self.net = nn.Linear(in_dim, self.out_dim)

    def forward(self, in_tensor: Tensor) -> Tensor:
        if self.net is None:
            raise SystemError("in_dim not set. Must be provided to constructor, or set_in_dim() should be called.")
        out = self.net(in_tensor)
        return self.activation(out) if self.activation else out


class DensityFieldHead(FieldHead):
    def __init__(self, in_dim: Optional[int] = None, activation: Optional[nn.Module] = nn.Softplus()) -> None:
        super().__init__(in_dim=in_dim, out_dim=1, field_head_name=FieldHeadNames.DENSITY, activation=activation)


class RGBFieldHead(FieldHead):
    def __init__(self, in_dim: Optional[int] = None, activation: Optional[nn.Module] = nn.Sigmoid()) -> None:
        super().__init__(in_dim=in_dim, out_dim=3, field_head_name=FieldHeadNames.RGB, activation=activation)
