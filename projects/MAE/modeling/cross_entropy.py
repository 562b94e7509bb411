"""Soft-target cross entropy for Mixup / CutMix fine-tuning (reference projects/MAE/modeling/cross_entropy.py)."""
from libai_b200.data.vision import SoftTargetCrossEntropy  # noqa: F401
