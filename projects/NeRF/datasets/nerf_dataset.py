"""Ray datasets for NeRF: Blender (synthetic 360°), LLFF (forward-facing / spherical real captures) and an analytic
scene for offline runs.

Spec: reference projects/NeRF/datasets/nerf_dataset.py:95-879 — camera-ray generation, NDC rays, LLFF pose
re-centring and spiral/spherical render paths, and the per-split sample conventions:

* ``train``  one item = ``batchsize`` random rays of one image (centre crop for the first ``precrop_iters`` items),
  fields ``rays`` [B, 8] and ``rgbs`` [B, 3];
* ``val`` / ``test``  one item = every ray of one image, plus ``c2w`` and ``valid_mask``;
* ``vis``  rays of a novel pose on the render path, plus ``c2w``.

The shared part (camera model → rays, sampling, item assembly) lives in ``RayDataset``; a concrete dataset only
supplies images, poses and bounds.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Optional

import numpy as np
import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance


# ------------------------------------------------------------------------------------------------ camera maths
def get_ray_directions(H, W, focal):
    """Per-pixel ray directions in camera coordinates (x right, y up, looking down -z): [H, W, 3]."""
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    return torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], dim=-1)


def get_rays(directions, c2w):
    """Rotate camera-space directions into world space; origin = camera centre.  Returns ([H*W,3], [H*W,3])."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    rays_d = directions @ c2w[:, :3].T
    rays_o = c2w[:, 3].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def get_ndc_rays(H, W, focal, near, rays_o, rays_d):
    """Project forward-facing rays into normalised device coordinates (NeRF paper, appendix C)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz, oy_oz = rays_o[..., 0] / rays_o[..., 2], rays_o[..., 1] / rays_o[..., 2]
    o = torch.stack([-focal / (W / 2) * ox_oz, -focal / (H / 2) * oy_oz, 1.0 + 2.0 * near / rays_o[..., 2]], -1)
    d = torch.stack([
        -focal / (W / 2) * (rays_d[..., 0] / rays_d[..., 2] - ox_oz),
        -focal / (H / 2) * (rays_d[..., 1] / rays_d[..., 2] - oy_oz),
        1.0 - o[..., 2],
    ], -1)
    return o, d


def normalize(v):
    return v / np.linalg.norm(v)


def viewmatrix(z, up, pos):
    z = normalize(z)
    x = normalize(np.cross(up, z))
    y = normalize(np.cross(z, x))
    return np.stack([x, y, z, pos], axis=1)


def average_poses(poses):
    """Mean camera: mean centre, mean viewing direction, mean up → [3, 4]."""
    return viewmatrix(poses[..., 2].mean(0), poses[..., 1].mean(0), poses[..., 3].mean(0))


def center_poses(poses):
    """Express ``poses`` [N,3,4] relative to their average pose.  Returns (centred poses, average pose 4×4)."""
    avg = np.eye(4)
    avg[:3] = average_poses(poses)
    homo = np.tile(np.eye(4), (len(poses), 1, 1))
    homo[:, :3] = poses
    return (np.linalg.inv(avg) @ homo)[:, :3], avg


def create_spiral_poses(radii, focus_depth, n_poses=120):
    """Spiral camera path for forward-facing scenes, all cameras looking at depth ``focus_depth``."""
    out = []
    for t in np.linspace(0, 4 * np.pi, n_poses + 1)[:-1]:
        center = np.array([np.cos(t), -np.sin(t), -np.sin(0.5 * t)]) * radii
        out.append(viewmatrix(center - np.array([0, 0, -focus_depth]), np.array([0, 1, 0]), center))
    return np.stack(out, 0)


def _rot_x(a):
    return np.array([[1, 0, 0, 0], [0, np.cos(a), -np.sin(a), 0], [0, np.sin(a), np.cos(a), 0], [0, 0, 0, 1.0]])


def _rot_y(a):
    return np.array([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1.0]])


def _trans_z(t):
    m = np.eye(4)
    m[2, 3] = t
    return m


def pose_spherical(theta, phi, radius):
    """Blender-convention camera on a sphere (degrees), looking at the origin: 4×4 tensor."""
    c2w = _rot_y(theta / 180.0 * np.pi) @ _rot_x(phi / 180.0 * np.pi) @ _trans_z(radius)
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]]) @ c2w
    return torch.tensor(c2w, dtype=torch.float32)


def create_spheric_poses(radius, n_poses=120):
    """Circular path around the up axis for 360° LLFF captures."""
    def one(theta, phi):
        m = np.eye(4)
        m[1, 3], m[2, 3] = -0.9 * radius, radius      # slightly above, at distance
        c2w = _rot_y(theta) @ _rot_x(phi) @ m
        return (np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]]) @ c2w)[:3]

    return np.stack([one(th, -np.pi / 5) for th in np.linspace(0, 2 * np.pi, n_poses + 1)[:-1]], 0)


# ------------------------------------------------------------------------------------------------ shared dataset
def to_instance(sample: dict) -> Instance:
    return Instance(**{k: DistTensorData(torch.as_tensor(v), placement_idx=0) for k, v in sample.items()})


class RayDataset(Dataset):
    """Subclasses fill: ``img_wh``, ``directions`` [H,W,3], ``train_poses``/``train_images`` (train split) or
    per-item ``_eval_view(idx)`` (other splits), and ``_bounds(rays_o)`` → (near, far)."""

    precrop_iters = 500
    precrop_frac = 0.5

    def __init__(self, root_dir, split, img_wh, batchsize=1024):
        self.root_dir, self.split, self.img_wh, self.batchsize = root_dir, split, tuple(img_wh), int(batchsize)
        self.num_iter = 0
        self._rng = np.random.default_rng(0)

    # -- ray assembly
    def rays_for_pose(self, c2w):
        rays_o, rays_d = get_rays(self.directions, torch.as_tensor(c2w, dtype=torch.float32)[:3, :4])
        rays_o, rays_d, near, far = self._finalize_rays(rays_o, rays_d)
        ones = torch.ones_like(rays_o[:, :1])
        return torch.cat([rays_o, rays_d, near * ones, far * ones], dim=1)

    def _finalize_rays(self, rays_o, rays_d):
        return rays_o, rays_d, self.near, self.far

    def build_train_bank(self, poses, images):
        """images: list of [H*W, 3] float tensors."""
        self.all_rays = torch.stack([self.rays_for_pose(p) for p in poses], 0)     # [N, H*W, 8]
        self.all_rgbs = torch.stack(list(images), 0)                              # [N, H*W, 3]

    def _train_item(self, idx):
        W, H = self.img_wh
        img = idx % self.all_rays.shape[0]
        if self.num_iter < self.precrop_iters:      # early iterations look at the image centre only
            dH, dW = int(H // 2 * self.precrop_frac), int(W // 2 * self.precrop_frac)
            ys = torch.arange(H // 2 - dH, H // 2 + dH)
            xs = torch.arange(W // 2 - dW, W // 2 + dW)
        else:
            ys, xs = torch.arange(H), torch.arange(W)
        pix = (ys[:, None] * W + xs[None, :]).reshape(-1)
        n = min(self.batchsize, pix.numel())
        sel = pix[torch.from_numpy(self._rng.choice(pix.numel(), size=n, replace=False))]
        self.num_iter += 1
        return dict(rays=self.all_rays[img, sel], rgbs=self.all_rgbs[img, sel])

    def __getitem__(self, idx):
        if self.split == "train":
            return to_instance(self._train_item(idx))
        return to_instance(self._eval_item(idx))

    def __len__(self):
        if self.split == "train":
            return int(self.all_rays.shape[0] * self.all_rays.shape[1] / self.batchsize)
        return self._eval_len()


NerfBaseDataset = RayDataset         # base-class name used by the reference (projects/NeRF/datasets/nerf_dataset.py:452)


def _load_image(path, img_wh, rgba):
    from PIL import Image

    img = Image.open(path)
    img = img.convert("RGBA" if rgba else "RGB").resize(tuple(img_wh), Image.LANCZOS)
    arr = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0)          # [H, W, C]
    return arr.reshape(-1, arr.shape[-1])


class BlenderDataset(RayDataset):
    """NeRF-synthetic scenes: ``transforms_{split}.json`` + RGBA PNGs, blended onto white."""

    def __init__(self, root_dir, split="train", img_wh=(800, 800), batchsize=1024, **kwargs):
        super().__init__(root_dir, split, img_wh, batchsize)
        self.white_back = True
        self.near, self.far = 2.0, 6.0
        self.bounds = np.array([self.near, self.far])
        meta_split = "train" if split == "vis" else split
        with open(os.path.join(root_dir, f"transforms_{meta_split}.json"), "r") as f:
            self.meta = json.load(f)
        w, h = self.img_wh
        self.focal = 0.5 * w / np.tan(0.5 * float(self.meta["camera_angle_x"]))
        self.directions = get_ray_directions(h, w, self.focal)
        self.render_poses = torch.stack([pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, 41)[:-1]], 0)
        if split == "train":
            poses, images = [], []
            for frame in self.meta["frames"]:
                poses.append(np.array(frame["transform_matrix"], dtype=np.float32)[:3, :4])
                images.append(self._blend(self._frame_image(frame))[0])
            self.build_train_bank(poses, images)

    def _frame_image(self, frame):
        return _load_image(os.path.join(self.root_dir, f"{frame['file_path']}.png"), self.img_wh, rgba=True)

    @staticmethod
    def _blend(rgba):
        return rgba[:, :3] * rgba[:, 3:] + (1.0 - rgba[:, 3:]), rgba[:, 3] > 0

    def _eval_len(self):
        if self.split == "val":
            return min(8, len(self.meta["frames"]))
        return len(self.render_poses) if self.split == "vis" else len(self.meta["frames"])

    def _eval_item(self, idx):
        if self.split == "vis":
            c2w = self.render_poses[idx][:3, :4]
            return dict(rays=self.rays_for_pose(c2w), c2w=c2w)
        frame = self.meta["frames"][idx]
        c2w = torch.tensor(frame["transform_matrix"], dtype=torch.float32)[:3, :4]
        rgb, valid = self._blend(self._frame_image(frame))
        return dict(rays=self.rays_for_pose(c2w), rgbs=rgb, c2w=c2w, valid_mask=valid)


class LLFFDataset(RayDataset):
    """Real captures with COLMAP poses (``poses_bounds.npy`` + ``images/``).  Forward-facing scenes are rendered in
    NDC space (near/far = 0/1); ``spheric_poses`` captures keep metric rays with bounds from the point cloud."""

    def __init__(self, root_dir, split="train", img_wh=(504, 378), spheric_poses=False, val_num=1, batchsize=1024,
                 **kwargs):
        super().__init__(root_dir, split, img_wh, batchsize)
        self.spheric_poses, self.val_num, self.white_back = bool(spheric_poses), max(1, int(val_num or 1)), False
        poses_bounds = np.load(os.path.join(root_dir, "poses_bounds.npy"))              # [N, 17]
        self.image_paths = sorted(glob.glob(os.path.join(root_dir, "images/*")))
        if split in ("train", "val"):
            assert len(poses_bounds) == len(self.image_paths), \
                "Mismatch between number of images and number of poses! Please rerun COLMAP!"
        poses = poses_bounds[:, :15].reshape(-1, 3, 5)
        self.bounds = poses_bounds[:, -2:].copy()
        H, W, focal = (float(x) for x in poses[0, :, -1])
        assert H * self.img_wh[0] == W * self.img_wh[1], \
            f"You must set @img_wh to have the same aspect ratio as ({W}, {H}) !"
        self.focal = focal * self.img_wh[0] / W
        # LLFF stores (down, right, back); NeRF wants (right, up, back)
        poses = np.concatenate([poses[..., 1:2], -poses[..., :1], poses[..., 2:4]], axis=-1)
        self.poses, self.pose_avg = center_poses(poses)
        self.val_idx = int(np.argmin(np.linalg.norm(self.poses[..., 3], axis=1)))       # most central image
        scale = self.bounds.min() * 0.75                                               # nearest depth → ~1.33
        self.bounds /= scale
        self.poses[..., 3] /= scale
        w, h = self.img_wh
        self.directions = get_ray_directions(h, w, self.focal)
        self.hwf = np.array([h, w, self.focal])
        if self.spheric_poses:
            self.near = float(self.bounds.min())
            self.far = float(min(8 * self.near, self.bounds.max()))                     # central object only
        else:
            self.near, self.far = 0.0, 1.0
        if split == "train":
            keep = [i for i in range(len(self.image_paths)) if i != self.val_idx]
            self.build_train_bank([self.poses[i] for i in keep], [self._image(self.image_paths[i]) for i in keep])
        elif split not in ("val",):
            if split.endswith("train"):
                self.poses_test = self.poses
            elif not self.spheric_poses:
                radii = np.percentile(np.abs(self.poses[..., 3]), 90, axis=0)
                self.poses_test = create_spiral_poses(radii, focus_depth=3.5)
            else:
                self.poses_test = create_spheric_poses(1.1 * self.bounds.min())

    def _image(self, path):
        rgb = _load_image(path, self.img_wh, rgba=False)
        return rgb

    def _finalize_rays(self, rays_o, rays_d):
        if not self.spheric_poses:
            rays_o, rays_d = get_ndc_rays(self.img_wh[1], self.img_wh[0], self.focal, 1.0, rays_o, rays_d)
        return rays_o, rays_d, self.near, self.far

    def _eval_len(self):
        return self.val_num if self.split == "val" else len(self.poses_test)

    def _eval_item(self, idx):
        if self.split == "val":
            c2w = torch.tensor(self.poses[self.val_idx], dtype=torch.float32)
            rgb = self._image(self.image_paths[self.val_idx])
            return dict(rays=self.rays_for_pose(c2w), rgbs=rgb, c2w=c2w,
                        valid_mask=torch.ones(rgb.shape[0], dtype=torch.bool))
        c2w = torch.tensor(self.poses_test[idx], dtype=torch.float32)
        return dict(rays=self.rays_for_pose(c2w), c2w=c2w)


class AnalyticSceneDataset(RayDataset):
    """A procedurally shaded sphere on white, rendered analytically from cameras on a ring — the Blender conventions
    without any files (offline smoke training, unit tests, benchmarks)."""

    def __init__(self, root_dir=None, split="train", img_wh=(32, 32), batchsize=256, n_views=8, **kwargs):
        super().__init__(root_dir, split, img_wh, batchsize)
        self.white_back, self.near, self.far = True, 2.0, 6.0
        w, h = self.img_wh
        self.focal = 0.5 * w / np.tan(0.5 * 0.6911)
        self.directions = get_ray_directions(h, w, self.focal)
        offset = {"train": 0.0, "val": 17.0, "test": 29.0, "vis": 41.0}[split]
        self.view_poses = [pose_spherical(a + offset, -30.0, 4.0) for a in np.linspace(-180, 180, n_views + 1)[:-1]]
        if split == "train":
            self.build_train_bank([p[:3, :4] for p in self.view_poses], [self._shade(p)[0] for p in self.view_poses])

    def _shade(self, c2w, radius=1.0):
        rays_o, rays_d = get_rays(self.directions, c2w[:3, :4])
        d = rays_d / rays_d.norm(dim=-1, keepdim=True)
        b = (rays_o * d).sum(-1)
        disc = b * b - ((rays_o * rays_o).sum(-1) - radius * radius)
        hit = disc > 0
        t = -b - disc.clamp(min=0).sqrt()
        normal = (rays_o + t[:, None] * d) / radius
        colour = 0.5 + 0.5 * normal                                                    # normal map shading
        rgb = torch.where(hit[:, None], colour, torch.ones_like(colour))
        return rgb, hit

    def _eval_len(self):
        return len(self.view_poses)

    def _eval_item(self, idx):
        c2w = self.view_poses[idx][:3, :4]
        if self.split == "vis":
            return dict(rays=self.rays_for_pose(c2w), c2w=c2w)
        rgb, hit = self._shade(self.view_poses[idx])
        return dict(rays=self.rays_for_pose(c2w), rgbs=rgb, c2w=c2w, valid_mask=hit)


def get_nerf_dataset(dataset_type="Blender"):
    table = {"Blender": BlenderDataset, "LLFF": LLFFDataset, "Analytic": AnalyticSceneDataset}
    assert dataset_type in table, f"The Nerf dataset must be one of {sorted(table)}"
    return table[dataset_type]
