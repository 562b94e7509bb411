"""Render a camera path from a trained NeRF into a video (reference
projects/NeRF/configs/config_nerf_for_rendering.py).

    bash tools/train.sh tools/train_net.py projects/NeRF/configs/config_nerf_for_rendering.py 1 --eval-only \
        train.load_weight=output/nerf/model_final
"""
from libai_b200.config import LazyCall
from libai_b200.data.build import build_image_test_loader
from projects.NeRF.configs.config_nerf import (  # noqa: F401
    _extra, _img_wh, _root, dataloader, dataset, graph, model, optim, train,
)
from projects.NeRF.evaluation.nerf_evaluator import NerfVisEvaluator

train.load_weight = "/path/to/checkpoint/model_final"
_split = "vis" if train.dataset_type != "LLFF" else "test"
_n_frames = {"Blender": 40, "Analytic": 8, "LLFF": 120}[train.dataset_type]
train.evaluation.evaluator = LazyCall(NerfVisEvaluator)(img_wh=_img_wh, pose_dir_len=_n_frames,
                                                        name=f"nerf_{train.dataset_type.lower()}_rendering",
                                                        image_save_path=train.output_dir)
dataloader.test = [
    LazyCall(build_image_test_loader)(
        dataset=LazyCall(dataset)(split=_split, img_wh=_img_wh, root_dir=_root, **_extra),
        num_workers=0,
        test_batch_size=1,
    )
]
