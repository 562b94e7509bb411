"""Attention helpers of the reference (projects/CLIP/clip/ops.py implements ``multi_head_attention_forward`` by hand);
here the fused op is ``libai_b200.ops.functional.attention``."""
from libai_b200.ops.functional import attention as multi_head_attention_forward  # noqa: F401
