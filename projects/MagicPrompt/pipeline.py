"""Prompt-completion pipeline (reference projects/MagicPrompt/pipeline.py)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from libai_b200.inference.basic import BasePipeline  # noqa: E402
from libai_b200.utils import distributed as dist  # noqa: E402


class TextGenerationPipeline(BasePipeline):
    def load_pretrain_weight(self, libai_cfg_model, model_path, mode="huggingface"):
        if mode == "huggingface":
            from libai_b200.models.utils.model_loader import GPT2LoaderHuggerFace

            loader = GPT2LoaderHuggerFace(libai_cfg_model, libai_cfg_model.cfg, model_path)
            loader.base_model_prefix_2 = ""          # the backbone is the model itself here
            model = loader.load()
            return model
        if mode == "libai":
            from libai_b200.models.utils.model_loader import GPT2LoaderLiBai

            return GPT2LoaderLiBai(libai_cfg_model, libai_cfg_model.cfg, model_path).load()
        return super().load_pretrain_weight(libai_cfg_model, model_path, mode=mode)

    def _parse_parameters(self, **pipeline_parameters):
        return {}, {**pipeline_parameters}, {}

    def preprocess(self, inputs, **kwargs) -> dict:
        ids = self.tokenizer.encode(inputs, return_tensors="pt")
        return {"input_ids": self.to_device(ids if ids.dim() == 2 else ids.unsqueeze(0))}

    def forward(self, inputs, **kwargs) -> dict:
        return {"return_ids": self.model.generate(inputs["input_ids"], **kwargs)}

    def postprocess(self, model_output_dict, **kwargs) -> dict:
        ids = model_output_dict["return_ids"]
        return [{"generated_text": self.tokenizer.decode(ids[i], skip_special_tokens=True)} for i in range(ids.shape[0])]


if __name__ == "__main__":
    pipeline = TextGenerationPipeline("projects/MagicPrompt/configs/gpt2_inference.py", data_parallel=1, tensor_parallel=1,
                                      pipeline_parallel=1, model_path="/path/to/magicprompt", mode="huggingface")
    out = pipeline(["a dog"], max_length=40)
    if dist.is_main_process():
        print(out)
