"""Import-path shim: the reference keeps this class in its own file (projects/MT5/layers/logits_layer.py); the implementation lives in libai_b200/layers/lm_logits.py."""
from libai_b200.layers import LMLogits  # noqa: F401
