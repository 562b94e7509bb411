"""Import-path shim: the reference keeps this class in its own file (projects/NeRF/modeling/System.py); the implementation lives in projects/NeRF/modeling/system.py."""
from projects.NeRF.modeling.system import NerfSystem  # noqa: F401
