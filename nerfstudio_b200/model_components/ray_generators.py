"""RayGenerator lives with the camera kernel; re-exported here under the reference's module path
(nerfstudio/model_components/ray_generators.py)."""
from ..cameras.cameras import RayGenerator  # noqa: F401
