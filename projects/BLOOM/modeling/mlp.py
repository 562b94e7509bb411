from projects.BLOOM.modeling.bloom_model import BloomMLP  # noqa: F401
