"""Datasets for Stable Diffusion fine-tuning (reference projects/Stable_Diffusion/dataset.py:28-188).

* ``DreamBoothDataset`` — a few *instance* images with one prompt, optionally paired with *class* (prior
  preservation) images and prompt; the two are concatenated along the batch so one forward covers both.
* ``PromptDataset`` — N copies of a prompt (+ index) for generating the class images on several GPUs.
* ``TXTDataset`` — a folder of ``xxx.jpg``/``xxx.txt`` pairs (LAION/COCO style captions).
"""
import os
from pathlib import Path

import torch
from torch.utils.data import Dataset
from torchvision import transforms

from libai_b200.data.structures import DistTensorData, Instance

_IMG_EXT = {".jpg", ".jpeg", ".png", ".bmp", ".webp"}


def _image_transform(size, center_crop):
    return transforms.Compose([
        transforms.Resize(size, interpolation=transforms.InterpolationMode.BILINEAR),
        transforms.CenterCrop(size) if center_crop else transforms.RandomCrop(size),
        transforms.ToTensor(),
        transforms.Normalize([0.5], [0.5]),
    ])


def _build_tokenizer(tokenizer, tokenizer_pretrained_folder):
    """``tokenizer`` is an instance or a class; with a class, ``tokenizer_pretrained_folder`` is ``path`` or
    ``[path, subfolder]``."""
    if isinstance(tokenizer, type):
        folder = tokenizer_pretrained_folder
        if isinstance(folder, (list, tuple)):
            return tokenizer.from_pretrained(folder[0], subfolder=folder[1])
        return tokenizer.from_pretrained(folder)
    return tokenizer


def _tokenize(tokenizer, text):
    ids = tokenizer(text, padding="max_length", truncation=True, max_length=tokenizer.model_max_length).input_ids
    return torch.tensor(ids, dtype=torch.long)


def _open(path):
    from PIL import Image

    img = Image.open(path)
    return img if img.mode == "RGB" else img.convert("RGB")


class DreamBoothDataset(Dataset):
    def __init__(self, instance_data_root, instance_prompt, tokenizer, tokenizer_pretrained_folder=None,
                 class_data_root=None, class_prompt=None, size=512, center_crop=False):
        self.tokenizer = _build_tokenizer(tokenizer, tokenizer_pretrained_folder)
        self.instance_images = sorted(p for p in Path(instance_data_root).iterdir() if p.suffix.lower() in _IMG_EXT)
        if not self.instance_images:
            raise ValueError(f"Instance images root {instance_data_root} has no images.")
        self.instance_prompt = instance_prompt
        self.class_images = None
        if class_data_root is not None:
            Path(class_data_root).mkdir(parents=True, exist_ok=True)
            self.class_images = sorted(p for p in Path(class_data_root).iterdir() if p.suffix.lower() in _IMG_EXT)
            self.class_prompt = class_prompt
        self._length = max(len(self.instance_images), len(self.class_images or []))
        self.transform = _image_transform(size, center_crop)

    def __len__(self):
        return self._length

    def __getitem__(self, index):
        pixel = self.transform(_open(self.instance_images[index % len(self.instance_images)]))
        ids = _tokenize(self.tokenizer, self.instance_prompt)
        if self.class_images:
            cls_pixel = self.transform(_open(self.class_images[index % len(self.class_images)]))
            cls_ids = _tokenize(self.tokenizer, self.class_prompt)
            # [2, ...]: the collate turns the batch into [B, 2, ...]; ``prior_preservation_collate`` below flattens
            # it to instance-half | class-half
            pixel, ids = torch.stack([pixel, cls_pixel]), torch.stack([ids, cls_ids])
        return Instance(pixel_values=DistTensorData(pixel), input_ids=DistTensorData(ids))


def prior_preservation_collate(batch):
    """[B × (2, C, H, W)] → [2B, C, H, W] with all instance samples first, then all class samples."""
    pix = torch.stack([b.get("pixel_values").tensor for b in batch])
    ids = torch.stack([b.get("input_ids").tensor for b in batch])
    if pix.dim() == 5:
        pix = pix.transpose(0, 1).reshape(-1, *pix.shape[2:])
        ids = ids.transpose(0, 1).reshape(-1, ids.shape[-1])
    return Instance(pixel_values=DistTensorData(pix), input_ids=DistTensorData(ids))


class PromptDataset(Dataset):
    def __init__(self, prompt, num_samples):
        self.prompt, self.num_samples = prompt, num_samples

    def __len__(self):
        return self.num_samples

    def __getitem__(self, index):
        return {"prompt": self.prompt, "index": index}


class TXTDataset(Dataset):
    def __init__(self, foloder_name, tokenizer, tokenizer_pretrained_folder=None, thres=0.2, size=512,
                 center_crop=False):
        self.tokenizer = _build_tokenizer(tokenizer, tokenizer_pretrained_folder)
        root = Path(foloder_name)
        self.image_paths = sorted(p for p in root.iterdir() if p.suffix.lower() in _IMG_EXT
                                  and p.with_suffix(".txt").exists())
        print(f"TXTDataset: {len(self.image_paths)} image/caption pairs in {root}")
        self.thres = thres          # kept for config compatibility (caption-dropout threshold of the reference)
        self.transform = _image_transform(size, center_crop)

    def __len__(self):
        return len(self.image_paths)

    def __getitem__(self, idx):
        path = self.image_paths[idx]
        with open(path.with_suffix(".txt"), "r", encoding="utf-8") as f:
            caption = f.read().strip()
        return Instance(pixel_values=DistTensorData(self.transform(_open(path))),
                        input_ids=DistTensorData(_tokenize(self.tokenizer, caption)))
