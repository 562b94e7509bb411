"""Greedy decoding of the second line of a couplet (reference projects/Couplets/infer.py)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import torch  # noqa: E402

from libai_b200.config import LazyConfig, instantiate  # noqa: E402
from libai_b200.utils import distributed as dist  # noqa: E402
from libai_b200.utils.checkpoint import Checkpointer  # noqa: E402
from projects.Couplets.dataset.mask import make_padding_mask, make_sequence_mask  # noqa: E402
from projects.Couplets.tokenizer.tokenizer import CoupletsTokenizer  # noqa: E402


class GeneratorForEager:
    def __init__(self, config_file, checkpoint_file, vocab_file):
        cfg = LazyConfig.load(config_file)
        dist.setup_dist_util(cfg.train.dist)
        self.model = instantiate(cfg.model).eval()
        if checkpoint_file:
            Checkpointer(self.model).load(checkpoint_file, checkpointables=[])
        self.tokenizer = CoupletsTokenizer(vocab_file)

    @torch.no_grad()
    def infer(self, sentence):
        t = self.tokenizer
        dev = next(self.model.parameters()).device
        enc = [t.bos_id] + t.convert_tokens_to_ids(t.tokenize(sentence)) + [t.eos_id]
        enc_ids = torch.tensor([enc], device=dev)
        enc_mask = torch.from_numpy(make_padding_mask(enc, enc, t.pad_id))[None].to(dev)
        states = self.model.encode(enc_ids, enc_mask)
        dec = [t.bos_id]
        for _ in range(len(enc) - 1):
            dec_mask = torch.from_numpy(make_padding_mask(dec, dec, t.pad_id) * make_sequence_mask(dec))[None].to(dev)
            cross = torch.from_numpy(make_padding_mask(dec, enc, t.pad_id))[None].to(dev)
            logits = self.model.decode(torch.tensor([dec], device=dev), dec_mask, states, cross)
            nxt = int(logits[0, -1].argmax())
            dec.append(nxt)
            if nxt == t.eos_id:
                break
        return "".join(t.convert_ids_to_tokens(dec[1:-1] if dec[-1] == t.eos_id else dec[1:]))


if __name__ == "__main__":
    gen = GeneratorForEager("projects/Couplets/configs/config.py", "output/couplet/model_final", "data_test/couplets/vocabs")
    print(gen.infer("天增岁月人增寿"))
