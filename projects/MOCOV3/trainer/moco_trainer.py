"""Trainer for MoCo v3 (reference projects/MOCOV3/trainer/moco_trainer.py).  The reference threads the iteration and
the momentum through ``run_step``; here the model owns both (``MoCo.cu_iter``), so the default step applies."""
from libai_b200.engine.trainer import StepTrainer


class MoCoEagerTrainer(StepTrainer):
    pass
