"""Import-path shim: the reference keeps this class in its own file (projects/NeRF/modeling/NeRF.py); the implementation lives in projects/NeRF/modeling/nerf.py."""
from projects.NeRF.modeling.nerf import Embedding, NeRF  # noqa: F401
