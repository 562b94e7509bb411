from .arguments import default_argument_parser
from .config import configurable, get_config, try_get_key
from .dictconfig import DictConfig, ListConfig, OmegaConf
from .instantiate import instantiate
from .lazy import LazyCall, LazyConfig

__all__ = [
    "LazyCall",
    "LazyConfig",
    "instantiate",
    "default_argument_parser",
    "configurable",
    "try_get_key",
    "get_config",
    "DictConfig",
    "ListConfig",
    "OmegaConf",
]
