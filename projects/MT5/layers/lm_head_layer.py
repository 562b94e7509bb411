"""Import-path shim: the reference keeps this class in its own file (projects/MT5/layers/lm_head_layer.py); the implementation lives in libai_b200/layers/linear.py."""
from libai_b200.layers import Linear


class LMHead(Linear):
    """Untied output projection ``hidden -> vocab`` (column-parallel, no bias) used by mT5 (T5 v1.1)."""

    def __init__(self, model_type, hidden_size, vocab_size, hidden_layers):
        super().__init__(hidden_size, vocab_size, bias=False, parallel="col", layer_idx=2 * hidden_layers - 1)
        self.model_type = model_type
