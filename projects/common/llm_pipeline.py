"""Causal-LM text generation pipeline shared by the Llama-family projects (reference projects/Llama/pipeline.py and
its copies): tokenise a prompt, ``model.generate``, decode.  ``mode``: ``"huggingface"`` (loader class supplied by the
project), ``"libai"``, ``"random"``."""
from libai_b200.inference.basic import BasePipeline


class CausalLMPipeline(BasePipeline):
    hf_loader = None      # set by the project
    libai_loader = None

    def load_pretrain_weight(self, libai_cfg_model, model_path, mode="huggingface"):
        if mode == "huggingface":
            return self.hf_loader(libai_cfg_model, libai_cfg_model.cfg, model_path).load()
        if mode == "libai" and self.libai_loader is not None:
            return self.libai_loader(libai_cfg_model, libai_cfg_model.cfg, model_path).load()
        return super().load_pretrain_weight(libai_cfg_model, model_path, mode=mode)

    def _parse_parameters(self, **pipeline_parameters):
        return {}, {**pipeline_parameters}, {}

    def preprocess(self, inputs, **kwargs) -> dict:
        ids = self.tokenizer.tokenize(inputs, add_bos=True, padding=True, device=None)
        return {"input_ids": self.to_device(ids)}

    def forward(self, inputs, **kwargs) -> dict:
        if "max_length" not in kwargs and "max_new_tokens" not in kwargs:
            kwargs["max_length"] = min(inputs["input_ids"].shape[1] + 64, self.cfg.model.cfg.max_position_embeddings)
        return {"return_ids": self.model.generate(inputs["input_ids"], **kwargs)}

    def postprocess(self, model_output_dict, **kwargs) -> dict:
        ids = model_output_dict["return_ids"]
        return [{"generated_text": self.tokenizer.decode(ids[i])} for i in range(ids.shape[0])]
