"""Import-path shim: the reference keeps this class in its own file (projects/NeRF/optimizers/Radam.py); the implementation lives in projects/NeRF/optimizers/radam.py."""
from projects.NeRF.optimizers.radam import RAdam  # noqa: F401
