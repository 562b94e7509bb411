"""Google-BERT style ``FullTokenizer`` API (reference projects/QQP/tokenizer/bert_tokenization.py) on the library's
WordPiece implementation."""
from libai_b200.tokenizer.tokenization_bert import (  # noqa: F401
    BasicTokenizer,
    BasicTokenizerWithChineseWWM,
    WordpieceTokenizer,
    load_vocab,
    whitespace_tokenize,
)


class FullTokenizer:
    def __init__(self, vocab_file, do_lower_case=True, do_chinese_wwm=False):
        self.vocab = load_vocab(vocab_file)
        self.inv_vocab = {v: k for k, v in self.vocab.items()}
        basic = BasicTokenizerWithChineseWWM if do_chinese_wwm else BasicTokenizer
        self.basic_tokenizer = basic(do_lower_case=do_lower_case)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab, unk_token="[UNK]")

    def tokenize(self, text):
        return [sub for tok in self.basic_tokenizer.tokenize(text) for sub in self.wordpiece_tokenizer.tokenize(tok)]

    def convert_tokens_to_ids(self, tokens):
        return [self.vocab.get(t, self.vocab["[UNK]"]) for t in tokens]

    def convert_ids_to_tokens(self, ids):
        return [self.inv_vocab[i] for i in ids]

    def vocab_size(self):
        return len(self.vocab)
