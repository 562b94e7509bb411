from libai_b200.config import LazyCall
from projects.MAE.modeling.mae import MaskedAutoencoderViT

model = LazyCall(MaskedAutoencoderViT)(
    img_size=224, patch_size=14, in_chans=3, embed_dim=1280, depth=32, num_heads=16,
    decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4, norm_pix_loss=True, mask_ratio=0.75,
)
