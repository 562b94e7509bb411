from .clip import available_models, load, tokenize  # noqa: F401
from .model import CLIP, build_model, convert_weights  # noqa: F401
