"""Shared helpers of the sentence-pair / span-corruption datasets.

Spec: reference libai/data/data_utils/dataset_utils.py — ``is_shared_folder`` (:35-49),
``compile_helper`` (:52-64), ``create_masked_lm_predictions`` (:79-322; n-gram / whole-word masking in
the styles ``bert``, ``bert-cn-wwm`` and ``t5``), ``get_samples_mapping`` (:325-413; cached ``.npy`` index
built once per node or once per shared folder), ``get_train_valid_test_split_`` (:416-444).
The numpy ``RandomState`` call sequence of the masking routine is part of the data-order contract
and is preserved; the code is organised differently (candidate building / n-gram drawing / token
replacement are separate helpers).
"""
from __future__ import annotations

import collections
import logging
import os
import re
import time

import numpy as np

from libai_b200.utils import distributed as dutil

logger = logging.getLogger(__name__)

MaskedLmInstance = collections.namedtuple("MaskedLmInstance", ["index", "label"])
_CN_SUBWORD = re.compile("##[一-龥]")


def is_shared_folder(filename) -> bool:
    """True when ``filename`` lives on storage shared between nodes (``IS_SHARED_FILE=1`` forces it)."""
    if os.environ.get("IS_SHARED_FILE") == "1":
        return True
    path = os.path.abspath(filename)
    if os.stat(path).st_dev == os.stat("/").st_dev:
        return False
    parent = os.path.dirname(path)
    return os.path.ismount(parent) and os.path.realpath(parent).startswith(("/nfs", "/smb", "/cifs"))


def compile_helper():
    """Build the native index helpers (single process!)."""
    from . import helpers_build

    helpers_build.ensure_built()


def is_start_piece(piece) -> bool:
    """WordPiece continuation pieces start with ``##``."""
    return not piece.startswith("##")


def _strip_cn_subword(tokenizer, token_id):
    tok = tokenizer.convert_ids_to_tokens([token_id])[0]
    if _CN_SUBWORD.findall(tok):
        tok = tok[2:]
    return tokenizer.convert_tokens_to_ids([tok])[0]


def _word_candidates(tokens, id_to_token, cls_id, sep_id, whole_word):
    """Group token positions into maskable units; returns (units, token_boundary flags)."""
    units, boundary = [], [0] * len(tokens)
    for i, tok in enumerate(tokens):
        if tok == cls_id or tok == sep_id:
            boundary[i] = 1
            continue
        starts = is_start_piece(id_to_token[tok])
        if whole_word and units and not starts:
            units[-1].append(i)
        else:
            units.append([i])
            if starts:
                boundary[i] = 1
    return units, boundary


def create_masked_lm_predictions(
    tokenizer,
    tokens,
    vocab_id_list,
    vocab_id_to_token_dict,
    masked_lm_prob,
    cls_id,
    sep_id,
    mask_id,
    max_predictions_per_seq,
    np_rng,
    max_ngrams=3,
    do_whole_word_mask=True,
    favor_longer_ngram=False,
    do_permutation=False,
    geometric_dist=False,
    masking_style="bert",
):
    """Masked-LM corruption of a token-id sequence.

    Returns ``(output_tokens, masked_positions, masked_labels, token_boundary, masked_spans)``
    (4-tuple without spans when ``masked_lm_prob == 0``, like the reference)."""
    if masking_style not in ("bert", "bert-cn-wwm", "t5"):
        raise ValueError("invalid value of masking style")
    units, token_boundary = _word_candidates(tokens, vocab_id_to_token_dict, cls_id, sep_id, do_whole_word_mask)
    output = list(tokens)
    if masking_style == "bert-cn-wwm":
        output = [_strip_cn_subword(tokenizer, t) for t in output]
    if masked_lm_prob == 0:
        return output, [], [], token_boundary

    budget = min(max_predictions_per_seq, max(1, int(round(len(tokens) * masked_lm_prob))))
    ngram_sizes = np.arange(1, max_ngrams + 1, dtype=np.int64)
    pvals = None
    if not geometric_dist:
        pvals = 1.0 / np.arange(1, max_ngrams + 1)
        pvals /= pvals.sum(keepdims=True)
        if favor_longer_ngram:
            pvals = pvals[::-1]
    # windows[i][n-1] = the n consecutive units starting at unit i
    windows = [[units[i : i + n] for n in ngram_sizes] for i in range(len(units))]
    np_rng.shuffle(windows)

    def draw_span(window, used, cap, rng):
        """Pick an n-gram from ``window`` that fits into ``cap - used`` positions (or None)."""
        if not geometric_dist:
            p = pvals[: len(window)]
            n = rng.choice(ngram_sizes[: len(window)], p=p / p.sum(keepdims=True))
        else:
            n = min(rng.geometric(0.2), max_ngrams)
        span = sum(window[n - 1], [])
        n -= 1
        while used + len(span) > cap and n > 0:
            span = sum(window[n - 1], [])
            n -= 1
        return None if used + len(span) > cap else span

    def replacement(pos):
        if masking_style == "t5":
            return mask_id
        if np_rng.random() < 0.8:
            return mask_id
        if np_rng.random() < 0.5:
            return tokens[pos] if masking_style == "bert" else _strip_cn_subword(tokenizer, tokens[pos])
        return vocab_id_list[np_rng.randint(0, len(vocab_id_list))]

    masked, spans, covered = [], [], set()
    for window in windows:
        if len(masked) >= budget:
            break
        if not window:
            continue
        span = draw_span(window, len(masked), budget, np_rng)
        if span is None or any(i in covered for i in span):
            continue
        for pos in span:
            covered.add(pos)
            output[pos] = replacement(pos)
            masked.append(MaskedLmInstance(index=pos, label=tokens[pos]))
        spans.append(MaskedLmInstance(index=span, label=[tokens[i] for i in span]))
    assert len(masked) <= budget

    np_rng.shuffle(windows)
    if do_permutation:
        chosen = set()
        for window in windows:
            if len(chosen) >= budget:
                break
            if not window:
                continue
            span = draw_span(window, len(chosen), budget, np.random)  # global RNG, as in the reference
            if span is None or any(i in covered or i in chosen for i in span):
                continue
            chosen.update(span)
        assert len(chosen) <= budget
        src = sorted(chosen)
        dst = list(src)
        np_rng.shuffle(dst)
        before = list(output)
        for s_i, t_i in zip(src, dst):
            output[s_i] = before[t_i]
            masked.append(MaskedLmInstance(index=s_i, label=before[s_i]))

    masked.sort(key=lambda m: m.index)
    spans.sort(key=lambda m: m.index[0])
    return output, [m.index for m in masked], [m.label for m in masked], token_boundary, spans


def get_samples_mapping(indexed_dataset, data_prefix, num_epochs, max_num_samples, max_seq_length, short_seq_prob,
                        seed, name, binary_head):
    """(start sentence, end sentence, target length) per sample; built by the C++ helper on one rank
    per node (or one rank overall on shared storage), cached next to the data as ``.npy``."""
    if not num_epochs:
        if not max_num_samples:
            raise ValueError("Need to specify either max_num_samples or num_epochs")
        num_epochs = np.iinfo(np.int32).max - 1
    if not max_num_samples:
        max_num_samples = np.iinfo(np.int64).max - 1
    fname = f"{data_prefix}_{name}_indexmap"
    if num_epochs != np.iinfo(np.int32).max - 1:
        fname += f"_{num_epochs}ep"
    if max_num_samples != np.iinfo(np.int64).max - 1:
        fname += f"_{max_num_samples}mns"
    fname += f"_{max_seq_length}msl_{short_seq_prob:0.2f}ssp_{seed}s.npy"

    folder = os.path.dirname(fname) or "."
    builder_rank = dutil.get_rank() if is_shared_folder(folder) else dutil.get_local_rank()
    if builder_rank == 0 and not os.path.isfile(fname):
        logger.info(f" > WARNING: could not find index map file {fname}, building the indices on rank 0 ...")
        assert indexed_dataset.doc_idx.dtype == np.int64
        assert indexed_dataset.sizes.dtype == np.int32
        from . import helpers

        t0 = time.time()
        logger.info(f" > building samples index mapping for {name} ...")
        mapping = helpers.build_mapping(
            indexed_dataset.doc_idx, indexed_dataset.sizes, num_epochs, max_num_samples, max_seq_length,
            short_seq_prob, seed, dutil.get_local_rank() == 0, 2 if binary_head else 1,
        )
        np.save(fname, mapping, allow_pickle=True)
        logger.info(f" > saved the index mapping in {fname} ({time.time() - t0:4f} s)")
    dutil.synchronize()
    logger.info(f" > loading indexed mapping from {fname}")
    t0 = time.time()
    mapping = np.load(fname, allow_pickle=True, mmap_mode="r")
    logger.info(f"    loaded indexed file in {time.time() - t0:3.3f} seconds; total number of samples: {mapping.shape[0]}")
    return mapping


def get_train_valid_test_split_(size, splits=None):
    """Boundaries ``[0, a, b, size]`` of the train/valid/test partitions for proportions ``splits``."""
    splits = list(splits) if splits is not None else [0.8, 0.2, 0.0]
    splits = (splits + [0.0, 0.0, 0.0])[:3]
    total = sum(splits)
    assert total > 0.0, "Split sum must be larger than 0."
    bounds = [0]
    for frac in splits:
        bounds.append(bounds[-1] + int(round(frac / total * float(size))))
    excess = bounds[-1] - size
    bounds = [bounds[0]] + [b - excess for b in bounds[1:]]
    assert len(bounds) == 4 and bounds[-1] == size
    return bounds
