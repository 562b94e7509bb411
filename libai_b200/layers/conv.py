"""GPT-2 style ``Conv1D``: a linear layer whose weight is stored ``[in, out]``.

Spec: reference libai/layers/conv.py:25-127 — same col/row semantics as ``Linear1D`` with the
split dimensions swapped (col splits weight dim 1, row splits weight dim 0).
"""
from torch import nn

from libai_b200.ops import functional as OF
from libai_b200.parallel import mappings
from libai_b200.utils import distributed as dutil

from ._param import create_parameter, xavier_normal_, zeros_


class Conv1D(nn.Module):
    def __init__(
        self,
        in_features,
        out_features,
        bias=True,
        parallel="data",
        init_method=xavier_normal_,
        skip_bias_add=False,
        dtype=None,
        *,
        layer_idx=0,
    ):
        super().__init__()
        if parallel not in ("data", "col", "row"):
            raise KeyError(f"{parallel} is not supported! Only support ('data', 'row' and 'col')")
        self.in_features, self.out_features = in_features, out_features
        self.parallel, self.skip_bias_add = parallel, skip_bias_add
        w_dim = {"data": None, "col": 1, "row": 0}[parallel]
        self.weight = create_parameter(
            (in_features, out_features), init_method, tp_dim=w_dim, layer_idx=layer_idx, dtype=dtype
        )
        self.bias = (
            create_parameter(
                (out_features,), zeros_, tp_dim=0 if parallel == "col" else None, layer_idx=layer_idx, dtype=dtype
            )
            if bias
            else None
        )
        if parallel == "row" and self.bias is not None:
            self.bias.sequence_parallel = dutil.get_dist_util().sequence_parallel

    def forward(self, x):
        sp = dutil.get_dist_util().sequence_parallel
        bias_now = None if self.skip_bias_add else self.bias
        wt = self.weight.t()  # [out, in] view; the GEMM front-end handles the layout
        if self.parallel == "col":
            x = mappings.gather_from_sp(x) if sp else mappings.copy_to_tp(x)
            y = OF.linear(x, wt.contiguous(), bias_now)
        elif self.parallel == "row":
            y = OF.linear(x, wt.contiguous(), None)
            y = mappings.reduce_scatter_to_sp(y) if sp else mappings.reduce_from_tp(y)
            if bias_now is not None:
                y = y + bias_now.to(y.dtype)
        else:
            y = OF.linear(x, wt.contiguous(), bias_now)
        return (y, self.bias) if self.skip_bias_add else y

    def extra_repr(self) -> str:
        return "in_features={}, out_features={}, bias={}, parallel={}".format(
            self.in_features, self.out_features, self.bias is not None, self.parallel
        )
