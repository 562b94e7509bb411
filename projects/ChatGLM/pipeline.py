"""ChatGLM generation / chat pipeline (reference projects/ChatGLM/pipeline.py)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from libai_b200.utils import distributed as dist  # noqa: E402
from projects.ChatGLM.utils.chatglm_loader import ChatGLMLoaderHuggerFace, ChatGLMLoaderLiBai  # noqa: E402
from projects.common.llm_pipeline import CausalLMPipeline  # noqa: E402


class TextGenerationPipeline(CausalLMPipeline):
    hf_loader = ChatGLMLoaderHuggerFace
    libai_loader = ChatGLMLoaderLiBai

    def preprocess(self, inputs, **kwargs) -> dict:
        return {"input_ids": self.to_device(self.tokenizer.tokenize(inputs, add_bos=True, padding=True))}

    def forward(self, inputs, **kwargs) -> dict:
        kwargs.setdefault("max_length", min(inputs["input_ids"].shape[1] + 128, self.cfg.model.cfg.seq_length))
        return {"return_ids": self.model.generate(inputs["input_ids"], **kwargs)}

    def chat(self, query, history=None, **kwargs):
        return self.model.chat(self.tokenizer, query, history=history, **kwargs)


if __name__ == "__main__":
    pipeline = TextGenerationPipeline("projects/ChatGLM/configs/chatglm_config.py", data_parallel=1, tensor_parallel=1,
                                      pipeline_parallel=1, model_path=os.getenv("CHATGLM_HF_DIR"), mode="huggingface")
    out = pipeline(["Give three tips for staying healthy."])
    if dist.is_main_process():
        print(out)
