"""The fused attention op vs ``torch.nn.functional.multi_head_attention_forward`` (reference
projects/CLIP/tests/test_multi_head_attn.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))


def test_mha_matches_torch():
    from projects.CLIP.clip.model import MultiheadAttention

    torch.manual_seed(0)
    ours = MultiheadAttention(32, 4).eval()
    ref = torch.nn.MultiheadAttention(32, 4).eval()
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(6, 2, 32)
    with torch.no_grad():
        a = ours(x, causal=True)
        mask = torch.full((6, 6), float("-inf")).triu_(1)
        b = ref(x, x, x, attn_mask=mask, need_weights=False)[0]
    assert (a - b).abs().max() < 1e-5
