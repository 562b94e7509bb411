"""Llama-7B shaped causal LM config (reference projects/Llama/configs/llama_config.py)."""
from libai_b200.config import DictConfig, LazyCall
from libai_b200.models import LlamaForCausalLM

cfg = DictConfig(
    dict(
        hidden_layers=32, vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_attention_heads=32,
        max_position_embeddings=2048, rms_norm_eps=1e-6, initializer_range=0.02,
        use_scaled_init_for_output_weights=True, scale_mask_softmax_fusion=False, amp_enabled=True,
        # generation defaults
        is_encoder_decoder=False, max_length=256, min_length=0, do_sample=False, early_stopping=False, num_beams=1,
        num_beam_groups=1, diversity_penalty=0.0, temperature=0.9, top_k=50, top_p=0.6, typical_p=1.0,
        repetition_penalty=1.0, length_penalty=1.0, no_repeat_ngram_size=0, encoder_no_repeat_ngram_size=0,
        num_return_sequences=1, chunk_size_feed_forward=0, output_scores=False, use_cache=True,
        bos_token_id=1, eos_token_id=2, pad_token_id=0, pretrained_model_path=None,
    )
)

model = LazyCall(LlamaForCausalLM)(cfg=cfg)
