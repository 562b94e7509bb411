"""NeRF model defaults (reference projects/NeRF/configs/config_model.py)."""
from libai_b200.config import DictConfig, LazyCall
from projects.NeRF.modeling.system import NerfSystem

cfg = DictConfig(dict(
    D=8, W=256, in_channels_xyz=63, in_channels_dir=27, skips=[4], N_samples=64, use_disp=False, perturb=1.0,
    noise_std=1.0, N_importance=128, chunk=64 * 1204, dataset_type="Blender", loss_func=None,
))

model = LazyCall(NerfSystem)(cfg=cfg)
