"""Distributed couplet generation on :class:`BasePipeline` (reference projects/Couplets/distribute_infer.py:23-118).

``CoupletPipeline(config, data_parallel, tensor_parallel, pipeline_parallel, …, model_path=…, vocab_file=…)`` builds the
topology, loads the seq2seq transformer and answers ``pipeline("天增岁月人增寿")`` with the greedy-decoded second line.
Launch with ``bash tools/infer.sh projects/Couplets/distribute_infer.py <gpus>``.
"""
import torch

from libai_b200.inference.basic import BasePipeline
from libai_b200.utils import distributed as dist
from projects.Couplets.dataset.mask import make_padding_mask, make_sequence_mask
from projects.Couplets.infer import GeneratorForEager  # noqa: F401  (single-process variant)
from projects.Couplets.tokenizer.tokenizer import CoupletsTokenizer


class CoupletPipeline(BasePipeline):
    def __init__(self, config_file, *args, vocab_file="data_test/couplets/vocab.txt", **kwargs):
        self.vocab_file = vocab_file
        super().__init__(config_file, *args, **kwargs)

    def _parse_parameters(self, **pipeline_parameters):
        return {}, {k: v for k, v in pipeline_parameters.items() if k == "max_extra_tokens"}, {}

    def load_pretrain_weight(self, libai_cfg_model, model_path, mode="libai"):
        if mode == "random":
            return super().load_pretrain_weight(libai_cfg_model, model_path, mode)
        from libai_b200.config import instantiate
        from libai_b200.utils.checkpoint import Checkpointer

        if "pretrained_model_path" in libai_cfg_model.cfg:      # set by BasePipeline; not a model hyper-parameter here
            del libai_cfg_model.cfg["pretrained_model_path"]
        model = instantiate(libai_cfg_model)
        Checkpointer(model).load(model_path, checkpointables=[])
        return model

    def build_tokenizer(self, cfg):
        return CoupletsTokenizer(self.vocab_file)

    def preprocess(self, sentence, **kwargs) -> dict:
        t = self.tokenizer
        enc = [t.bos_id] + t.convert_tokens_to_ids(t.tokenize(sentence)) + [t.eos_id]
        return {"encoder_ids": enc}

    def forward(self, inputs, max_extra_tokens: int = 0, **kwargs) -> dict:
        t, enc = self.tokenizer, inputs["encoder_ids"]
        dev = next(self.model.parameters()).device
        enc_ids = torch.tensor([enc], device=dev)
        enc_mask = torch.from_numpy(make_padding_mask(enc, enc, t.pad_id))[None].to(dev)
        states = self.model.encode(enc_ids, enc_mask)
        dec = [t.bos_id]
        for _ in range(len(enc) - 1 + max_extra_tokens):     # a couplet's second line has the length of the first
            dec_mask = torch.from_numpy(make_padding_mask(dec, dec, t.pad_id) * make_sequence_mask(dec))[None].to(dev)
            cross = torch.from_numpy(make_padding_mask(dec, enc, t.pad_id))[None].to(dev)
            logits = self.model.decode(torch.tensor([dec], device=dev), dec_mask, states, cross)
            nxt = int(logits[0, -1].argmax())
            dec.append(nxt)
            if nxt == t.eos_id:
                break
        return {"return_ids": torch.tensor(dec)}

    def postprocess(self, outputs, **kwargs) -> dict:
        ids = outputs["return_ids"].tolist()
        ids = ids[1:-1] if ids and ids[-1] == self.tokenizer.eos_id else ids[1:]
        return {"generated_text": "".join(self.tokenizer.convert_ids_to_tokens(ids))}

    # name used by the reference's script
    def generate(self, sentence):
        return self(sentence).get("generated_text")


if __name__ == "__main__":
    import sys

    pipeline = CoupletPipeline("projects/Couplets/configs/config.py", data_parallel=1, tensor_parallel=1, pipeline_parallel=1,
                               model_path="output/couplet/model_final", vocab_file="data_test/couplets/vocab.txt")
    out = pipeline(sys.argv[1] if len(sys.argv) > 1 else "天增岁月人增寿")
    if dist.is_main_process():
        print(out)
