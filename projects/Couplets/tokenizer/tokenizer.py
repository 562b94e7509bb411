"""Character-level tokenizer for Chinese couplets (reference projects/Couplets/tokenizer/tokenizer.py): one token
per character, ``<pad> <unk> <bos> <eos>`` at the head of the vocabulary."""
import collections


def load_vocab(vocab_file):
    vocab = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as f:
        for i, line in enumerate(f):
            vocab[line.rstrip("\n")] = i
    return vocab


class CoupletsTokenizer:
    def __init__(self, vocab_file):
        self.vocab = load_vocab(vocab_file)
        self.inv_vocab = {v: k for k, v in self.vocab.items()}
        self.pad_id, self.unk_id = self.vocab.get("<pad>", 0), self.vocab.get("<unk>", 1)
        self.bos_id, self.eos_id = self.vocab.get("<bos>", 2), self.vocab.get("<eos>", 3)

    def tokenize(self, text):
        return text.split() if " " in text.strip() else list(text.strip())

    def convert_tokens_to_ids(self, tokens_list):
        return [self.vocab.get(t, self.unk_id) for t in tokens_list]

    def convert_ids_to_tokens(self, ids_list):
        return [self.inv_vocab.get(int(i), "<unk>") for i in ids_list]
