"""User API: ``available_models``, ``load``, ``tokenize`` (reference projects/CLIP/clip/clip.py).  Checkpoints are the
OpenAI TorchScript / state-dict files; there is no network here, so ``load`` takes a local path (or a name that
``libai_b200.utils.file_utils.cached_path`` can resolve from a warm cache)."""
import os
from typing import List, Union

import torch
from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor

from .model import build_model
from .simple_tokenizer import SimpleTokenizer

_MODELS = {
    "RN50": "https://openaipublic.azureedge.net/clip/models/afeb0e10f9e5a86da6080e35cf09123aca3b358a0c3e3b6c78a7b63bc04b6762/RN50.pt",
    "RN101": "https://openaipublic.azureedge.net/clip/models/8fa8567bab74a42d41c5915025a8e4538c3bdbe8804a470a72f30b0d94fab599/RN101.pt",
    "ViT-B/32": "https://openaipublic.azureedge.net/clip/models/40d365715913c9da98579312b702a82c18be219cc2a73407c4526f58eba950af/ViT-B-32.pt",
    "ViT-B/16": "https://openaipublic.azureedge.net/clip/models/5806e77cd80f8b59890b7e101eabd078d9fb84e6937f9e85e4ecb61988df416f/ViT-B-16.pt",
    "ViT-L/14": "https://openaipublic.azureedge.net/clip/models/b8cca3fd41ae0c99ba7e8951adf17d267cdb84cd88be6f7c2e0eca1737a03836/ViT-L-14.pt",
}
_tokenizer = None


def available_models() -> List[str]:
    return list(_MODELS.keys())


def _convert_image_to_rgb(image):
    return image.convert("RGB")


def _transform(n_px):
    return Compose([Resize(n_px, interpolation=InterpolationMode.BICUBIC), CenterCrop(n_px), _convert_image_to_rgb, ToTensor(),
                    Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


def load(name: str, device: Union[str, torch.device] = None, download_root: str = None):
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    if name in _MODELS:
        from libai_b200.utils.file_utils import cached_path

        path = cached_path(_MODELS[name], cache_dir=download_root or os.path.expanduser("~/.cache/clip"))
    elif os.path.isfile(name):
        path = name
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    try:
        state_dict = torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        state_dict = torch.load(path, map_location="cpu", weights_only=False)
        state_dict = state_dict.get("state_dict", state_dict)
    model = build_model(dict(state_dict)).to(device)
    if str(device) == "cpu":
        model.float()
    return model, _transform(model.visual.input_resolution)


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False, bpe_path: str = None):
    global _tokenizer
    if _tokenizer is None:
        _tokenizer = SimpleTokenizer(bpe_path)
    texts = [texts] if isinstance(texts, str) else texts
    sot, eot = _tokenizer.encoder["<|startoftext|>"], _tokenizer.encoder["<|endoftext|>"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        tokens = [sot] + _tokenizer.encode(text) + [eot]
        if len(tokens) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            tokens = tokens[:context_length]
            tokens[-1] = eot
        result[i, : len(tokens)] = torch.tensor(tokens)
    return result
