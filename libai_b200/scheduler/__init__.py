from .build import build_lr_scheduler
from .lr_scheduler import (
    ClosedFormLR,
    WarmupCosineAnnealingLR,
    WarmupCosineLR,
    WarmupExponentialLR,
    WarmupMultiStepLR,
    WarmupPolynomialLR,
    WarmupStepLR,
)
