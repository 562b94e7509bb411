from libai_b200.config import LazyCall
from projects.MAE.modeling.vit import VisionTransformer

model = LazyCall(VisionTransformer)(
    img_size=224, patch_size=14, in_chans=3, embed_dim=1280, depth=32, num_heads=16, mlp_ratio=4,
    drop_path_rate=0.1, global_pool=True, num_classes=1000, loss_func=None,
)
