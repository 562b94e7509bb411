"""Qwen2 tokenizer (reference projects/Qwen/tokenizer.py): byte-level BPE with the Qwen pre-tokenisation regex
and the ChatML control tokens."""
from projects.common.sft import ByteBPEChatTokenizer

PRETOKENIZE_REGEX = (
    r"""(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"""
)


class Qwen2Tokenizer(ByteBPEChatTokenizer):
    def __init__(self, vocab_file, merges_file, bos_token=None, eos_token="<|endoftext|>", pad_token="<|endoftext|>",
                 unk_token="<|endoftext|>", **kwargs):
        super().__init__(vocab_file, merges_file, bos_token=bos_token, eos_token=eos_token, pad_token=pad_token,
                         unk_token=unk_token, special_tokens=["<|endoftext|>", "<|im_start|>", "<|im_end|>"],
                         pattern=PRETOKENIZE_REGEX, **kwargs)
