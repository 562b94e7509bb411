"""Megatron-style BERT tokenizer wrapper used by the QQP example (reference projects/QQP/tokenizer/tokenizer.py:
``_BertCNWWMTokenizer`` with ``cls/sep/pad/mask`` id properties, extra ids and BOS/EOS tokens)."""
from libai_b200.tokenizer import BertTokenizer


class _BertCNWWMTokenizer(BertTokenizer):
    def __init__(self, vocab_file, lower_case=True, vocab_extra_ids=0, **kwargs):
        extra = [f"<extra_id_{i}>" for i in range(vocab_extra_ids)]
        super().__init__(vocab_file, do_lower_case=lower_case, bos_token="[BOS]", eos_token="[EOS]",
                         additional_special_tokens=extra or None, **kwargs)
        self.sanitize_special_tokens()

    def add_token(self, token):
        self.add_tokens([token], special_tokens=True)
        return self.convert_tokens_to_ids(token)

    def add_additional_special_tokens(self, tokens_list):
        self.add_special_tokens({"additional_special_tokens": list(self.additional_special_tokens) + list(tokens_list)})

    @property
    def inv_vocab(self):
        return {v: k for k, v in self.get_vocab().items()}

    def decode_token_ids(self, token_ids):
        tokens = [t for t in self.convert_ids_to_tokens(token_ids) if t not in ("[PAD]", "[CLS]")]
        return " ".join(tokens).replace(" ##", "")

    cls = property(lambda self: self.cls_token_id)
    sep = property(lambda self: self.sep_token_id)
    pad = property(lambda self: self.pad_token_id)
    mask = property(lambda self: self.mask_token_id)
    bos_token_id_ = property(lambda self: self.convert_tokens_to_ids("[BOS]"))
    eos_token_id_ = property(lambda self: self.convert_tokens_to_ids("[EOS]"))
