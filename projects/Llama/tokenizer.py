"""``LlamaTokenizer`` (reference projects/Llama/tokenizer.py): sentencepiece model + bos/eos/pad handling."""
from projects.common.sft import SentencePieceTokenizer


class LlamaTokenizer(SentencePieceTokenizer):
    pass
