from libai_b200.config import LazyCall
from projects.MOCOV3.modeling.moco import MoCo_ViT
from projects.MOCOV3.modeling.vit import VisionTransformer

_vit = dict(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, drop_path_rate=0.0, stop_grad_conv1=True)
base_encoder = LazyCall(VisionTransformer)(**_vit)
momentum_encoder = LazyCall(VisionTransformer)(**_vit)

model = LazyCall(MoCo_ViT)(base_encoder=base_encoder, momentum_encoder=momentum_encoder, dim=256, mlp_dim=4096, T=0.2,
                           m=0.99, max_iter=300)
