"""Aquila tokenizer (reference projects/Aquila/tokenizer.py): byte-level BPE with ``<|startofpiece|>`` /
``<|endofpiece|>`` / ``<|endoftext|>`` style control tokens appended to the vocabulary."""
from projects.common.sft import ByteBPEChatTokenizer


class AquilaTokenizer(ByteBPEChatTokenizer):
    def __init__(self, vocab_file, merges_file, bos_token="<|startofpiece|>", eos_token="<|endofpiece|>",
                 pad_token="<|endoftext|>", unk_token="<|endoftext|>", **kwargs):
        specials = ["<|endoftext|>", "[UNK]", "[CLS]", "[SEP]", "[gMASK]", "<|startofpiece|>", "<|endofpiece|>"]
        super().__init__(vocab_file, merges_file, bos_token=bos_token, eos_token=eos_token, pad_token=pad_token,
                         unk_token=unk_token, special_tokens=specials, **kwargs)
