"""Baichuan tokenizer (reference projects/Baichuan/tokenizer.py): sentencepiece model."""
from projects.common.sft import SentencePieceTokenizer


class BaichuanTokenizer(SentencePieceTokenizer):
    pass
