"""ImageNet without labels for self-supervised pre-training (reference projects/MAE/data/pretraining_imagenet.py)."""
from libai_b200.data.datasets import ImageNetDataset
from libai_b200.data.structures import DistTensorData, Instance


class PretrainingImageNetDataset(ImageNetDataset):
    def __getitem__(self, index):
        inst = super().__getitem__(index)
        return Instance(images=DistTensorData(inst.get("images").tensor, placement_idx=0))
