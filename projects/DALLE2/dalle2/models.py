"""Single import point for the DALL·E 2 model classes.

The reference keeps everything in one 3 000-line ``dalle2/models.py``; here the implementation is split by role —
``diffusion.py`` (noise schedules), ``prior.py`` (causal transformer prior), ``unet.py`` + ``decoder.py`` (cascaded
decoder), ``dalle2.py`` (text → image driver) — and this module re-exports the public names so
``from projects.DALLE2.dalle2.models import DiffusionPrior, Decoder, Unet, DALLE2`` keeps working.
"""
from projects.DALLE2.dalle2.clip_adapter import OpenAIClipAdapter  # noqa: F401
from projects.DALLE2.dalle2.dalle2 import DALLE2  # noqa: F401
from projects.DALLE2.dalle2.decoder import Decoder, LowresConditioner  # noqa: F401
from projects.DALLE2.dalle2.diffusion import NoiseScheduler  # noqa: F401
from projects.DALLE2.dalle2.prior import (  # noqa: F401
    Attention,
    CausalTransformer,
    DiffusionPrior,
    DiffusionPriorNetwork,
    FeedForward,
    RelPosBias,
    SinusoidalPosEmb,
    SwiGLU,
)
from projects.DALLE2.dalle2.unet import (  # noqa: F401
    Block,
    ChanLayerNorm,
    CrossAttention,
    CrossEmbedLayer,
    ResnetBlock,
    Unet,
)
