"""Text cleaning and pair encoding for the GLUE tsv files (reference projects/QQP/dataset/data_utils.py)."""
import re

import numpy as np


def clean_text(text):
    text = text.replace("\n", " ")
    text = re.sub(r"\s+", " ", text)
    for _ in range(3):
        text = text.replace(" . ", ". ")
    return text


def build_tokens_types_paddings_from_text(text_a, text_b, tokenizer, max_seq_length):
    a = tokenizer.convert_tokens_to_ids(tokenizer.tokenize(text_a))
    b = tokenizer.convert_tokens_to_ids(tokenizer.tokenize(text_b)) if text_b is not None else None
    return build_tokens_types_paddings_from_ids(a, b, max_seq_length, tokenizer.cls_token_id, tokenizer.sep_token_id,
                                                tokenizer.pad_token_id)


def build_tokens_types_paddings_from_ids(a, b, max_seq_length, cls_id, sep_id, pad_id):
    ids, types = [cls_id] + list(a) + [sep_id], [0] * (len(a) + 2)
    if b is not None:
        ids += list(b)
        types += [1] * len(b)
    trimmed = len(ids) >= max_seq_length
    if trimmed:
        ids, types = ids[: max_seq_length - 1], types[: max_seq_length - 1]
    if b is not None or trimmed:
        ids.append(sep_id)
        types.append(1 if b is not None else 0)
    mask = [1] * len(ids)
    pad = max_seq_length - len(ids)
    return ids + [pad_id] * pad, types + [pad_id] * pad, mask + [0] * pad


def build_sample(ids, types, paddings, label, unique_id):
    return {"text": np.array(ids, dtype=np.int64), "types": np.array(types, dtype=np.int64),
            "padding_mask": np.array(paddings, dtype=np.int64), "label": int(label), "uid": int(unique_id)}
