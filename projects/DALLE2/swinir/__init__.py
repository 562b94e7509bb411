from .models import SwinIR  # noqa: F401
from .upsample import load_model, upsample4x, upsample16x  # noqa: F401
