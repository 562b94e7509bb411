from projects.text_classification.modeling.load_megatron_weight import convert_megatron_state, load_megatron_bert  # noqa: F401
