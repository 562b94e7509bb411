from .clip_adapter import OpenAIClipAdapter  # noqa: F401
from .dalle2 import DALLE2  # noqa: F401
from .decoder import Decoder  # noqa: F401
from .prior import DiffusionPrior, DiffusionPriorNetwork  # noqa: F401
from .unet import Unet  # noqa: F401
