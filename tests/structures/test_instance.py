"""Instance / DistTensorData containers (model: reference tests/structures/{test_instance,test_metadata}.py)."""
import pytest
import torch

from libai_b200.data.structures import DistTensorData, Instance


def test_init_args_and_fields():
    inst = Instance(images=torch.rand(4, 5))
    inst.tokens = torch.rand(4, 5, 6)
    assert inst.has("images") and inst.has("tokens")
    inst.remove("images")
    assert not inst.has("images")
    inst.meta_tensor = DistTensorData(torch.rand(5, 6))
    assert inst.has("meta_tensor") and isinstance(inst.get("meta_tensor"), DistTensorData)
    assert set(inst.get_fields()) == {"tokens", "meta_tensor"}
    with pytest.raises(AttributeError):
        inst.missing_field


def test_order_and_stack():
    a = Instance(tokens=DistTensorData(torch.arange(6).view(2, 3)), labels=DistTensorData(torch.tensor(1), placement_idx=-1))
    b = Instance(tokens=DistTensorData(torch.arange(6).view(2, 3) + 6), labels=DistTensorData(torch.tensor(0), placement_idx=-1))
    assert list(a.get_fields()) == ["tokens", "labels"]  # insertion order is the argument order of model.forward
    batch = Instance.stack([a, b])
    assert tuple(batch.tokens.tensor.shape) == (2, 2, 3) and batch.labels.tensor.tolist() == [1, 0]
    assert batch.labels.placement_idx == -1


def test_stack_rejects_mismatched_metadata():
    a = DistTensorData(torch.zeros(2), placement_idx=0)
    b = DistTensorData(torch.zeros(2), placement_idx=-1)
    with pytest.raises(AssertionError):
        DistTensorData.stack([a, b])


def test_to_global_moves_to_device():
    t = DistTensorData(torch.arange(4))
    t.to_global(device_type="cpu")
    assert t.tensor.device.type == "cpu" and t.tensor.tolist() == [0, 1, 2, 3]
