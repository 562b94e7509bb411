"""PSNR evaluation and image/video dumping for NeRF.

Spec: reference projects/NeRF/evaluation/nerf_evaluator.py:18-162 — per validation image: PSNR of the rendered
``rgb_{fine|coarse}`` against ground truth (optionally only over ``valid_mask``), optionally saving the prediction,
ground truth and a colour-mapped depth image; ``NerfVisEvaluator`` collects novel-view frames into a video/GIF.
OpenCV/imageio are optional: without them images are written with PIL and frames are kept in memory / saved as GIF.
"""
import copy
import os
import time
from collections import OrderedDict

import numpy as np
import torch

from libai_b200.evaluation.evaluator import DatasetEvaluator
from libai_b200.utils import distributed as dist


def mse(image_pred, image_gt, valid_mask=None, reduction="mean"):
    value = (image_pred - image_gt) ** 2
    if valid_mask is not None:
        value = value[valid_mask]
    return value.mean() if reduction == "mean" else value


def psnr(image_pred, image_gt, valid_mask=None, reduction="mean"):
    return -10.0 * torch.log10(mse(image_pred, image_gt, valid_mask, reduction))


def visualize_depth(depth):
    """[H, W] depth → [3, H, W] float image, normalised over the finite values and mapped blue→red."""
    x = np.nan_to_num(depth.detach().float().cpu().numpy())
    lo, hi = float(x.min()), float(x.max())
    x = (x - lo) / (hi - lo + 1e-8)
    try:
        import cv2

        img = cv2.applyColorMap((x * 255).astype(np.uint8), cv2.COLORMAP_JET)[..., ::-1] / 255.0
    except ImportError:
        img = np.stack([np.clip(1.5 - np.abs(4 * x - 3), 0, 1), np.clip(1.5 - np.abs(4 * x - 2), 0, 1),
                        np.clip(1.5 - np.abs(4 * x - 1), 0, 1)], axis=-1)
    return torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).float()


def _save_image(chw, path):
    from PIL import Image

    arr = (chw.clamp(0, 1).permute(1, 2, 0).cpu().numpy() * 255).astype(np.uint8)
    Image.fromarray(arr).save(path)


class NerfEvaluator(DatasetEvaluator):
    def __init__(self, img_wh, image_save_path=None):
        self.img_wh = tuple(img_wh)
        self.image_save_path = image_save_path
        if image_save_path is not None:
            os.makedirs(image_save_path, exist_ok=True)
        self._predictions = []
        self._toimage_count = 0

    def reset(self):
        self._predictions = []

    @staticmethod
    def current_time():
        return time.strftime("%Y%m%d-%H%M%S", time.localtime())

    def process(self, inputs, outputs):
        typ = "fine" if "rgb_fine" in outputs else "coarse"
        pred = outputs[f"rgb_{typ}"].float().reshape(-1, 3)
        gt = outputs["rgbs"].float().reshape(-1, 3).to(pred.device)
        mask = inputs.get("valid_mask", None) if hasattr(inputs, "get") else None
        value = psnr(pred, gt).item()
        self._predictions.append({"psnr": value, "losses": float(outputs.get("losses", 0.0))})
        if self.image_save_path is not None and dist.is_main_process():
            W, H = self.img_wh
            stem = os.path.join(self.image_save_path, f"{self.current_time()}_{self._toimage_count:03d}")
            _save_image(pred.view(H, W, 3).permute(2, 0, 1), stem + "_pred.png")
            _save_image(gt.view(H, W, 3).permute(2, 0, 1), stem + "_gt.png")
            _save_image(visualize_depth(outputs[f"depth_{typ}"].reshape(H, W)), stem + "_depth.png")
            self._toimage_count += 1

    def evaluate(self):
        if not dist.is_main_process():
            return {}
        n = max(1, len(self._predictions))
        self._results = OrderedDict(psnr=sum(p["psnr"] for p in self._predictions) / n)
        return copy.deepcopy(self._results)


class NerfVisEvaluator(NerfEvaluator):
    """Collect the rendered frames of a camera path and write ``<name>.gif`` (or ``.mp4`` with imageio-ffmpeg)."""

    def __init__(self, img_wh, pose_dir_len, name, image_save_path="."):
        super().__init__(img_wh, image_save_path=None)
        self.pose_dir_len, self.name, self.out_dir = pose_dir_len, name, image_save_path
        self.frames = []

    @staticmethod
    def to8b(x):
        return (255 * np.clip(x, 0, 1)).astype(np.uint8)

    def process(self, inputs, outputs):
        typ = "fine" if "rgb_fine" in outputs else "coarse"
        W, H = self.img_wh
        self.frames.append(self.to8b(outputs[f"rgb_{typ}"].float().reshape(H, W, 3).cpu().numpy()))
        if len(self.frames) == self.pose_dir_len and dist.is_main_process():
            self.write()

    def write(self):
        os.makedirs(self.out_dir, exist_ok=True)
        try:
            import imageio

            imageio.mimwrite(os.path.join(self.out_dir, f"{self.name}.mp4"), self.frames, fps=30, quality=8)
        except Exception:
            from PIL import Image

            imgs = [Image.fromarray(f) for f in self.frames]
            imgs[0].save(os.path.join(self.out_dir, f"{self.name}.gif"), save_all=True, append_images=imgs[1:],
                         duration=33, loop=0)

    def evaluate(self):
        return {"frames": len(self.frames)} if dist.is_main_process() else {}
