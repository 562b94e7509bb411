"""The reference keeps cache-aware copies of the attention / transformer layers here; the library layers of
``libai_b200`` already take ``past_key_value`` / ``use_cache``."""
from libai_b200.layers import MultiheadAttention, TransformerLayer  # noqa: F401
