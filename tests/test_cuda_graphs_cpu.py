"""CUDA-graph capture is guarded: on CPU / unsupported layouts the trainer stays eager and says why."""
import torch
from torch import nn

from libai_b200.engine import cuda_graphs


class _Block(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.lin = nn.Linear(8, 8)
        self.drop = nn.Dropout(p)
        self.layer_idx = 0

    def forward(self, x):
        return self.drop(self.lin(x))


def test_cpu_guard_and_micro_batch_slots():
    # CPU tensors are never captured (dropout no longer excludes a block: the native kernels draw graph-safe Philox state)
    assert cuda_graphs.graph_transformer_blocks([_Block(0.1)], torch.randn(2, 8)) is None
    cuda_graphs.set_micro_batch_slot(5)
    assert cuda_graphs._CURRENT_SLOT == 5
    cuda_graphs.set_micro_batch_slot(0)


def test_launch_count_wrapper_keeps_module_identity():
    from libai_b200 import ops

    blk = _Block()
    n0 = ops.launch_count()
    same = cuda_graphs._with_launch_count(blk, 3, 5)
    assert same is blk and list(blk.state_dict()) == ["lin.weight", "lin.bias"]
    blk.train()
    blk(torch.randn(2, 8))
    assert ops.launch_count() - n0 == 8
    with torch.no_grad():
        blk(torch.randn(2, 8))
    assert ops.launch_count() - n0 == 11


def test_block_call_spec_classification():
    """Blocks may take model buffers (RoPE tables) and a right-padded key mask besides the hidden state."""
    from libai_b200.layers.attention import KeyPaddingMask

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.register_buffer("cos", torch.ones(4, 2), persistent=False)

    m = M()
    spec = cuda_graphs.classify_block_call(m, (), {"cos_cached": m.cos, "past": None})
    assert spec is not None and not spec.trivial and spec.key_lengths is None
    assert spec.matches((), {"cos_cached": m.cos}) and not spec.matches((), {"cos_cached": m.cos.clone()})
    assert cuda_graphs.classify_block_call(m, (), {"x": torch.ones(2)}) is None          # not a buffer: varies per call
    assert cuda_graphs.classify_block_call(m, (None,), {}).trivial

    right = KeyPaddingMask(torch.tensor([[1, 1, 0], [1, 0, 0]]))
    holes = KeyPaddingMask(torch.tensor([[1, 0, 1], [1, 0, 0]]))
    spec = cuda_graphs.classify_block_call(m, (right,), {})
    assert spec is not None and spec.key_lengths.tolist() == [2, 1]
    assert spec.matches((right,), {}) and not spec.matches((holes,), {}) and not spec.matches((), {})
    assert spec.graph_inputs((right,))[0] is right.lengths
    assert cuda_graphs.classify_block_call(m, (holes,), {}) is None
    again = KeyPaddingMask.from_lengths(right.lengths)
    assert again.is_prefix() and again.lengths is right.lengths
