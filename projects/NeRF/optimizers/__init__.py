from .radam import RAdam  # noqa: F401
from .ranger import Ranger  # noqa: F401
