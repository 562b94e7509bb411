"""CLIP adapters (file name of the reference: projects/DALLE2/dalle2/_clip.py); implementation in ``clip_adapter.py``."""
from projects.DALLE2.dalle2.clip_adapter import OpenAIClipAdapter  # noqa: F401

# one adapter implementation here: the abstract base of the reference and the OpenAI adapter coincide
BaseClipAdapter = OpenAIClipAdapter
