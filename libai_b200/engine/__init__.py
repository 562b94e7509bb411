from .default import DefaultTrainer, default_setup
from .trainer import EagerTrainer, GraphTrainer, HookBase, StepTrainer, TrainerBase

__all__ = ["DefaultTrainer", "default_setup", "HookBase", "TrainerBase", "StepTrainer", "EagerTrainer", "GraphTrainer"]
