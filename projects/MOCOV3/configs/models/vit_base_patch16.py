from libai_b200.config import LazyCall
from projects.MOCOV3.modeling.vit import VisionTransformer

model = LazyCall(VisionTransformer)(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12,
                                    drop_path_rate=0.1, num_classes=1000, linear_prob=True, loss_func=None)
