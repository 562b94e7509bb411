"""Evaluation config (reference projects/Eval_LLM/config.py): the parallel layout and what to evaluate."""
from libai_b200.config import DictConfig

parallel_config = DictConfig(
    dict(
        data_parallel_size=1,
        tensor_parallel_size=1,
        pipeline_parallel_size=1,
        pipeline_num_layers=32,
        device_type="cuda",
    )
)

eval_config = DictConfig(
    dict(
        pretrained_model_path="",
        hf_tokenizer_path="",
        model_type="llama",             # a key of special_arguments.json
        model_weight_type="libai",      # "libai" | "huggingface" | "random"
        # lm-evaluation-harness task names/globs when `lm_eval` is installed, and/or paths to local ``*.jsonl``
        # task files (see eval_harness.LocalTask) which need no extra package
        eval_tasks=["lambada_openai", "gsm8k"],
        batch_size_per_gpu=1,
        limit=None,
        save_filepath=None,
    )
)
