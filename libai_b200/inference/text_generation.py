"""Text generation pipeline (reference libai/inference/text_generation.py:22-110; there bound to the MT5 project
loaders).  ``mode="huggingface"`` picks the loader from the model class: T5/MT5 or Llama-family."""
from libai_b200.inference.basic import BasePipeline
from libai_b200.utils import distributed as dist


class TextGenerationPipeline(BasePipeline):
    def load_pretrain_weight(self, libai_cfg_model, model_path, mode="huggingface"):
        if mode == "huggingface":
            target = str(libai_cfg_model.get("_target_", ""))
            if "llama" in target.lower():
                from libai_b200.models.utils.model_loader.llama_loader import LlamaLoaderHuggerFace as Loader

                return Loader(libai_cfg_model, libai_cfg_model.cfg, model_path).load()
            from projects.MT5.utils.mt5_loader import T5LoaderHuggerFace

            return T5LoaderHuggerFace(libai_cfg_model, libai_cfg_model.cfg, model_path, hidden_dropout_prob=0.0,
                                      attention_probs_dropout_prob=0.0, embedding_dropout_prob=0.0).load()
        return super().load_pretrain_weight(libai_cfg_model, model_path, mode=mode)

    def _parse_parameters(self, **pipeline_parameters):
        return {}, {**pipeline_parameters}, {}

    def preprocess(self, inputs, pad: bool = False, **kwargs) -> dict:
        encoder_ids = self.tokenizer.encode(inputs, return_tensors="pt")
        if encoder_ids.dim() == 1:
            encoder_ids = encoder_ids.unsqueeze(0)
        return {"encoder_ids": self.to_device(encoder_ids)}

    def forward(self, encoder_input_dict, **kwargs) -> dict:
        return {"return_ids": self.model.generate(encoder_input_dict["encoder_ids"], **kwargs)}

    def postprocess(self, model_output_dict, **kwargs) -> dict:
        ids = model_output_dict["return_ids"]
        return [{"generated_text": self.tokenizer.decode(ids[i], skip_special_tokens=True)} for i in range(ids.shape[0])]


if __name__ == "__main__":
    pipeline = TextGenerationPipeline(
        "projects/MT5/configs/t5_inference.py", data_parallel=1, tensor_parallel=2, pipeline_parallel=2,
        pipeline_stage_id=[0] * 12 + [1] * 12, pipeline_num_layers=12 * 2, model_path="/path/to/t5-base",
        mode="huggingface",
    )
    out = pipeline(["summarize: She is a student, She is tall, She loves study"])
    if dist.is_main_process():
        print(out)
