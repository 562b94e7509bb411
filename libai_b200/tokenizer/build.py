"""``build_tokenizer(cfg)`` — reference libai/tokenizer/build.py:23-33: instantiate ``cfg.tokenizer``; with
``append_eod`` make sure an end-of-document token exists (eos, else pad)."""
import logging

from libai_b200.config import instantiate

logger = logging.getLogger(__name__)


def build_tokenizer(cfg):
    tokenizer = instantiate(cfg.tokenizer)
    if cfg.get("append_eod", None) and tokenizer.eod_token is None:
        tokenizer.eod_token = tokenizer.eos_token if tokenizer.eos_token is not None else tokenizer.pad_token
    return tokenizer
