from . import distributed
