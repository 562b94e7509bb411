"""ChatGLM SFT entry point (reference projects/ChatGLM/train_net.py): HF weights → (optional) LoRA injection → train."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from libai_b200.config import LazyConfig, default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer, default_setup  # noqa: E402


class ChatGLMTrainer(DefaultTrainer):
    @classmethod
    def build_model(cls, cfg):
        path = try_get_key(cfg, "model.cfg.pretrained_model_path")
        if path and os.path.isdir(path):
            from projects.ChatGLM.utils.chatglm_loader import ChatGLMLoaderHuggerFace

            lora = cfg.model.cfg.lora_enable
            cfg.model.cfg.lora_enable = False          # load the dense weights first, then wrap
            model = ChatGLMLoaderHuggerFace(cfg.model, cfg.model.cfg, path).load()
            if lora:
                from projects.ChatGLM.lora.lora_model import LoraModel

                cfg.model.cfg.lora_enable = True
                model.transformer = LoraModel(model.transformer, cfg.model.cfg.lora_cfg, "default")
            return model
        return super().build_model(cfg)


def main(args):
    cfg = LazyConfig.apply_overrides(LazyConfig.load(args.config_file), args.opts)
    default_setup(cfg, args)
    return ChatGLMTrainer(cfg).train()


if __name__ == "__main__":
    main(default_argument_parser().parse_args())
