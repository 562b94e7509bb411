"""Tokenizers: common contract + per-tokenizer behaviour (model: reference tests/tokenizer/*).  The BERT tokenizer
is cross-checked against ``transformers``' slow implementation on the same vocabulary."""
import json
import os

import pytest

from libai_b200.tokenizer import BertTokenizer, GPT2Tokenizer, RobertaTokenizer, T5Tokenizer
from libai_b200.tokenizer.tokenization_gpt2 import bytes_to_unicode

BERT_VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "quick", "brown", "fox", "jump", "##s", "##ed", "over",
              "lazy", "dog", ",", ".", "un", "##want", "runn", "##ing", "中", "国", "人", "low", "##er", "##est"]
MERGES = [("t", "h"), ("th", "e"), ("Ġ", "the"), ("Ġ", "q"), ("i", "n"), ("in", "g"), ("Ġ", "d"), ("Ġd", "o"), ("Ġdo", "g"),
          ("e", "r"), ("l", "o"), ("lo", "w"), ("Ġ", "low"), ("Ġlow", "er")]


@pytest.fixture()
def bert_vocab(tmp_path):
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(BERT_VOCAB) + "\n", encoding="utf-8")
    return str(p)


@pytest.fixture()
def bpe_files(tmp_path):
    chars = list(bytes_to_unicode().values())
    vocab = {c: i for i, c in enumerate(chars)}
    for a, b in MERGES:
        vocab[a + b] = len(vocab)
    for tok in ["<|endoftext|>", "<s>", "</s>", "<unk>", "<pad>", "<mask>"]:
        vocab[tok] = len(vocab)
    v, m = tmp_path / "vocab.json", tmp_path / "merges.txt"
    v.write_text(json.dumps(vocab), encoding="utf-8")
    m.write_text("#version: 0.2\n" + "\n".join(f"{a} {b}" for a, b in MERGES) + "\n", encoding="utf-8")
    return str(v), str(m)


# ------------------------------------------------------------------ common contract
def _common(tok, text):
    ids = tok.encode(text)
    assert ids == tok.convert_tokens_to_ids(tok.tokenize(text)) or hasattr(tok, "build_inputs_with_special_tokens")
    assert tok.convert_ids_to_tokens(tok.convert_tokens_to_ids(tok.tokenize(text))) == tok.tokenize(text)
    n = len(tok)
    assert tok.add_tokens(["new_tok_aaa", "new_tok_bbb"]) == 2 and len(tok) == n + 2
    assert tok.add_tokens(["new_tok_aaa"]) == 0
    toks = tok.tokenize(f"{text} new_tok_aaa {text}")
    assert "new_tok_aaa" in toks
    assert tok.add_special_tokens({"additional_special_tokens": ["<special_x>"]}) == 1
    assert "<special_x>" in tok.all_special_tokens and tok.tokenize("a <special_x> b").count("<special_x>") == 1
    assert tok.padded_vocab_size(128) % 128 == 0 and tok.padded_vocab_size(128) >= len(tok)
    assert tok.encode([text, text], return_tensors="pt").shape[0] == 2


def test_bert_common_and_roundtrip(bert_vocab, tmp_path):
    tok = BertTokenizer(bert_vocab)
    _common(tok, "the quick brown fox")
    tok.save_pretrained(str(tmp_path / "saved"))
    tok2 = BertTokenizer.from_pretrained(str(tmp_path / "saved"))
    assert tok2.tokenize("the new_tok_aaa dog <special_x>") == tok.tokenize("the new_tok_aaa dog <special_x>")
    assert tok2.get_vocab() == tok.get_vocab()


def test_bert_full_tokenizer(bert_vocab):
    tok = BertTokenizer(bert_vocab, add_bos_token=True)
    assert tok.tokenize("UNwantéd,running") == ["un", "##want", "##ed", ",", "runn", "##ing"]
    assert tok.convert_tokens_to_ids(tok.tokenize("UNwantéd,running")) == [17, 18, 11, 15, 19, 20]
    assert tok.tokenize("ah博推zz") == ["[UNK]", "[UNK]", "[UNK]", "[UNK]"]
    assert tok.tokenize("lowest lower") == ["low", "##est", "low", "##er"]
    ids = tok.encode("the dog", )
    assert ids[0] == tok.cls_token_id and ids[-1] == tok.sep_token_id
    assert tok.build_inputs_with_special_tokens([5], [14]) == [2, 5, 3, 14, 3]
    assert tok.decode(tok.encode("the quick fox jumps."), skip_special_tokens=True) == "the quick fox jumps."
    assert tok.start_token == "[CLS]" and tok.end_token == "[SEP]" and tok.eod_token is None


def test_bert_matches_transformers(bert_vocab):
    hf = pytest.importorskip("transformers").BertTokenizer(bert_vocab)
    tok = BertTokenizer(bert_vocab, add_bos_token=True)
    for text in ["The quick brown fox jumps over the lazy dog.", "unwanted, running 中国人 [MASK] xyz", "Héllo   thE\tdog"]:
        assert tok.tokenize(text) == hf.tokenize(text)
        assert tok.encode(text) == hf.encode(text)


def test_bert_chinese_wwm(bert_vocab):
    tok = BertTokenizer(bert_vocab, do_chinese_wwm=True, pre_tokenizer=lambda x: [x[:2], x[2:]] if len(x) > 2 else [x])
    toks = tok.tokenize("中国人")
    assert toks == ["中", "##国", "人"]
    ids = tok.convert_tokens_to_ids(toks)
    assert ids[1] == tok.vocab_size + tok.vocab["国"] and tok.convert_ids_to_tokens(ids) == toks


def test_gpt2(bpe_files):
    tok = GPT2Tokenizer(*bpe_files)
    _common(tok, "the lower dog")
    tok = GPT2Tokenizer(*bpe_files, add_bos_token=True)
    text = "the quick dog is running — naïve café 中文!"
    ids = tok.encode(text)
    assert ids[0] == tok.bos_token_id and tok.decode(ids[1:]) == text          # byte-level: lossless
    assert tok.tokenize(" lower") == ["Ġlower"] and tok.tokenize("the dog") == ["the", "Ġdog"]
    assert tok.bpe("lower") == "low er"


def test_roberta(bpe_files):
    tok = RobertaTokenizer(*bpe_files, add_bos_token=True)
    ids = tok.encode("the dog <mask> running")
    toks = tok.convert_ids_to_tokens(ids)
    assert toks[0] == "<s>" and toks[-1] == "</s>" and "<mask>" in toks
    assert tok.create_token_type_ids_from_sequences([1, 2], [3]) == [0] * 7
    assert tok.build_inputs_with_special_tokens([7], [8]) == [tok.cls_token_id, 7, tok.sep_token_id, 8, tok.sep_token_id]


def test_t5(tmp_path):
    spm = pytest.importorskip("sentencepiece")
    corpus = tmp_path / "c.txt"
    corpus.write_text("\n".join(["the quick brown fox jumps over the lazy dog", "hello world this is a test",
                                 "machine translation with transformers"] * 50))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tmp_path / "spiece"), vocab_size=40, pad_id=0,
                                   eos_id=1, unk_id=2, bos_id=-1, minloglevel=2)
    tok = T5Tokenizer(str(tmp_path / "spiece.model"), add_bos_token=True)
    assert tok.vocab_size == 140
    assert tok.convert_tokens_to_ids("<extra_id_0>") == 139 and tok.convert_ids_to_tokens(40) == "<extra_id_99>"
    ids = tok.encode("the quick <extra_id_0> fox")
    assert ids[-1] == tok.eos_token_id and 139 in ids
    assert "the quick" in tok.decode(ids, skip_special_tokens=True)
    tok.save_pretrained(str(tmp_path / "saved"))
    assert os.path.exists(tmp_path / "saved" / "spiece.model")
    assert T5Tokenizer.from_pretrained(str(tmp_path / "saved"), add_bos_token=True).encode("the quick <extra_id_0> fox") == ids
