from .graph_base import GraphBase
from .weight_init import init_method_normal, scaled_init_method_normal
