"""Official MoCo v3 (timm ViT naming) → this project's names (reference projects/MOCOV3/utils/weight_convert.py)."""
from projects.MAE.utils.weight_convert import convert_state_dict  # noqa: F401  (same timm → libai_b200 mapping)
