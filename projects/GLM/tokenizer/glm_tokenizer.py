"""GLM tokenizers (reference projects/GLM/tokenizer/glm_tokenizer.py): a mixin adding the blank-infilling control
tokens and the generation / multiple-choice input builders on top of the library tokenizers."""
from typing import List, Optional

import torch

from libai_b200.tokenizer import BertTokenizer, GPT2Tokenizer, RobertaTokenizer
from libai_b200.tokenizer.tokenization_base import PreTrainedTokenizer


class GLMTokenizerMixin:
    sop_token = "<|startofpiece|>"
    eop_token = "<|endofpiece|>"

    @property
    def sop_token_id(self):
        return self.convert_tokens_to_ids(self.sop_token)

    @property
    def eop_token_id(self):
        return self.convert_tokens_to_ids(self.eop_token)

    @property
    def gmask_token_id(self):
        return self.convert_tokens_to_ids("[gMASK]")

    @property
    def smask_token_id(self):
        return self.convert_tokens_to_ids("[sMASK]")

    @property
    def mask_token_ids(self):
        return [self.mask_token_id, self.smask_token_id, self.gmask_token_id]

    def _register_glm_tokens(self):
        self.add_special_tokens({"additional_special_tokens": list(self.additional_special_tokens) +
                                 ["<|startofpiece|>", "<|endofpiece|>", "[gMASK]", "[sMASK]"]})

    def __call__(self, text, padding=True, return_tensors="pt", **kwargs):
        texts = [text] if isinstance(text, str) else list(text)
        rows = [self.encode(t) for t in texts]
        width = max(len(r) for r in rows)
        pad = self.pad_token_id if self.pad_token_id is not None else 0
        ids = torch.tensor([r + [pad] * (width - len(r)) for r in rows], dtype=torch.long)
        mask = torch.tensor([[1] * len(r) + [0] * (width - len(r)) for r in rows], dtype=torch.long)
        return {"input_ids": ids, "attention_mask": mask}

    def build_inputs_for_generation(self, model_input, max_gen_length=512, targets=None, padding=False):
        """Append ``<|startofpiece|>`` and lay out 2-D positions: generated tokens sit at the (first) mask position
        with block positions 1…n; the generation mask lets them see the whole context and their causal past."""
        input_ids = model_input["input_ids"]
        b, s = input_ids.shape[:2]
        labels = None
        if targets is not None:
            batched = isinstance(targets, (list, tuple))
            enc = [self.encode(t) for t in (targets if batched else [targets])]
            enc = [(t + [self.eop_token_id])[:max_gen_length] for t in enc]
            if not padding:
                max_gen_length = max(map(len, enc))
            enc = [[self.sop_token_id] + t for t in enc]
            labels = [t[1:] + [-100] * (max_gen_length - len(t) + 1) for t in enc]
            enc = [t + [self.pad_token_id] * (max_gen_length + 1 - len(t)) for t in enc]
            targets_t = torch.tensor(enc, dtype=input_ids.dtype)
            labels = torch.cat((input_ids.new_full((b, s), -100), torch.tensor(labels, dtype=input_ids.dtype)), dim=1)
        pos, blk = [], []
        for i in range(b):
            where = [p for m in self.mask_token_ids for p in (input_ids[i] == m).nonzero(as_tuple=True)[0].tolist()]
            if not where:
                raise ValueError("Cannot find mask token in the input")
            mask_pos = min(where)
            pos.append(list(range(s)) + [mask_pos] * max_gen_length)
            blk.append([0] * s + list(range(1, max_gen_length + 1)))
        position_ids = torch.stack((torch.tensor(pos), torch.tensor(blk)), dim=1).to(input_ids.dtype)
        ctx = model_input["attention_mask"].unsqueeze(1).expand(-1, s + max_gen_length, -1)
        gen = torch.cat([ctx.new_zeros((s, max_gen_length)), torch.tril(ctx.new_ones((max_gen_length, max_gen_length)))], 0)
        attention_mask = torch.cat((ctx, gen.unsqueeze(0).expand(b, -1, -1)), dim=2).unsqueeze(1)
        if targets is None:
            input_ids = torch.cat((input_ids, input_ids.new_full((b, 1), self.sop_token_id)), dim=-1)
        else:
            input_ids = torch.cat((input_ids, targets_t[:, :-1]), dim=1)
        batch = {"input_ids": input_ids, "position_ids": position_ids}
        if labels is None:
            batch["generation_attention_mask"] = attention_mask
        else:
            batch["attention_mask"], batch["labels"] = attention_mask, labels
        return batch

    def build_inputs_for_multiple_choice(self, model_input, choices, max_length=None):
        samples = []
        for i in range(len(model_input["input_ids"])):
            ctx_ids = model_input["input_ids"][i].tolist()
            ctx_mask = model_input["attention_mask"][i]
            division = len(ctx_ids)
            mask_position = ctx_ids.index(self.mask_token_id)
            token = torch.tensor(ctx_ids, dtype=torch.long)
            blocks = [ctx_mask.expand(division, -1)]
            position_id, block_position_id = torch.arange(division), torch.zeros(division, dtype=torch.long)
            choice_ids, choice_indices = [], []
            for text in choices[i]:
                ch = torch.tensor(self.encode(text), dtype=torch.long)
                choice_ids.append(ch)
                choice_indices.append(torch.arange(len(token), len(token) + len(ch)))
                blocks.append(torch.tril(torch.ones((len(ch), len(ch)), dtype=torch.long)))
                token = torch.cat((token, torch.tensor([self.sop_token_id]), ch[:-1]))
                position_id = torch.cat((position_id, torch.full((len(ch),), mask_position)))
                block_position_id = torch.cat((block_position_id, torch.arange(1, 1 + len(ch))))
            mask = torch.block_diag(*blocks)
            mask[division:, :division] = ctx_mask.unsqueeze(0)
            samples.append(dict(input_ids=token, position_ids=torch.stack((position_id, block_position_id)),
                                attention_mask=mask, choice_ids=choice_ids, choice_indices=choice_indices))
        width = max(len(x["input_ids"]) for x in samples)
        out = dict(input_ids=[], position_ids=[], attention_mask=[], choice_ids=[], choice_indices=[])
        for x in samples:
            pad = width - len(x["input_ids"])
            out["input_ids"].append(torch.cat((x["input_ids"], torch.zeros(pad, dtype=torch.long))))
            out["position_ids"].append(torch.cat((x["position_ids"], x["position_ids"][..., -1:].expand(-1, pad)), dim=-1))
            out["attention_mask"].append(torch.nn.functional.pad(x["attention_mask"], (0, pad, 0, pad)))
            out["choice_ids"].append(x["choice_ids"])
            out["choice_indices"].append(x["choice_indices"])
        return {"input_ids": torch.stack(out["input_ids"]), "position_ids": torch.stack(out["position_ids"]),
                "attention_mask": torch.stack(out["attention_mask"]).unsqueeze(1), "choice_ids": out["choice_ids"],
                "choice_indices": out["choice_indices"]}


class GLMRobertaTokenizer(GLMTokenizerMixin, RobertaTokenizer):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._register_glm_tokens()


class GLMGPT2Tokenizer(GLMTokenizerMixin, GPT2Tokenizer):
    def __init__(self, *args, cls_token="[CLS]", mask_token="[MASK]", pad_token="<|endoftext|>", **kwargs):
        super().__init__(*args, cls_token=cls_token, mask_token=mask_token, pad_token=pad_token, **kwargs)
        self._register_glm_tokens()

    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None):
        assert token_ids_1 is None
        return [self.cls_token_id] + token_ids_0 + [self.eos_token_id]


class GLMBertTokenizer(GLMTokenizerMixin, BertTokenizer):
    def __init__(self, *args, eos_token="[SEP]", **kwargs):
        kwargs.setdefault("add_bos_token", True)
        super().__init__(*args, eos_token=eos_token, **kwargs)
        self._register_glm_tokens()


class GLMChineseTokenzier(GLMTokenizerMixin, PreTrainedTokenizer):
    """sentencepiece tokenizer of the Chinese GLM checkpoints (class name keeps the reference's spelling)."""

    vocab_files_names = {"vocab_file": "cog-pretrain.model"}

    def __init__(self, vocab_file, eos_token="<|endoftext|>", unk_token="[UNK]", pad_token="<|endoftext|>",
                 cls_token="[CLS]", mask_token="[MASK]", **kwargs):
        import sentencepiece as spm

        super().__init__(eos_token=eos_token, unk_token=unk_token, pad_token=pad_token, cls_token=cls_token,
                         mask_token=mask_token, **kwargs)
        self.vocab_file = vocab_file
        self.sp_model = spm.SentencePieceProcessor()
        self.sp_model.Load(vocab_file)
        self._register_glm_tokens()

    @property
    def vocab_size(self):
        return len(self.sp_model)

    def get_vocab(self):
        vocab = {self.convert_ids_to_tokens(i): i for i in range(self.vocab_size)}
        vocab.update(self.added_tokens_encoder)
        return vocab

    def _tokenize(self, text, **kwargs):
        return self.sp_model.encode(text, out_type=str)

    def _convert_token_to_id(self, token):
        return self.sp_model.PieceToId(token)

    def _convert_id_to_token(self, index):
        return self.sp_model.IdToPiece(index)

    def convert_tokens_to_string(self, tokens):
        return self.sp_model.decode(tokens)

    def build_inputs_with_special_tokens(self, token_ids_0, token_ids_1=None):
        assert token_ids_1 is None
        return [self.cls_token_id] + token_ids_0 + [self.eos_token_id]

    def save_vocabulary(self, save_directory, filename_prefix=None):
        import os
        from shutil import copyfile

        out = os.path.join(save_directory, (filename_prefix + "-" if filename_prefix else "") + "cog-pretrain.model")
        if os.path.abspath(self.vocab_file) != os.path.abspath(out):
            copyfile(self.vocab_file, out)
        return (out,)
