"""Import-path shim: the reference keeps this class in its own file (projects/NeRF/optimizers/Ranger.py); the implementation lives in projects/NeRF/optimizers/ranger.py."""
from projects.NeRF.optimizers.ranger import Ranger  # noqa: F401
