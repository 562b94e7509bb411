"""Generate the class ("prior preservation") images DreamBooth regularises with.

Spec: reference projects/Stable_Diffusion/generate_prior_image.py — sample ``num_class_images`` images of
``class_prompt`` with the *un-tuned* model into ``class_data_dir``, sharding the work over the launched processes.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        projects/Stable_Diffusion/generate_prior_image.py --model_path CompVis/stable-diffusion-v1-4 \
        --class_prompt "a photo of dog" --class_data_dir /data/prior_dog --num_class_images 200
"""
import argparse
import hashlib
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from projects.Stable_Diffusion.dataset import PromptDataset  # noqa: E402
from projects.Stable_Diffusion.pipeline import StableDiffusionPipeline  # noqa: E402


def generate(pipeline, class_prompt, class_data_dir, num_class_images, batch_size=4, rank=0, world=1, **sample_kwargs):
    out = Path(class_data_dir)
    out.mkdir(parents=True, exist_ok=True)
    have = len([p for p in out.iterdir() if p.suffix in (".jpg", ".png")])
    if have >= num_class_images:
        return 0
    todo = PromptDataset(class_prompt, num_class_images - have)
    made = 0
    for start in range(rank * batch_size, len(todo), world * batch_size):
        items = [todo[i] for i in range(start, min(start + batch_size, len(todo)))]
        images = pipeline([it["prompt"] for it in items], **sample_kwargs)
        for it, image in zip(items, images):
            digest = hashlib.sha1(image.tobytes()).hexdigest()
            image.save(out / f"{have + it['index']}-{digest}.jpg")
            made += 1
    return made


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_path", required=True)
    ap.add_argument("--class_prompt", required=True)
    ap.add_argument("--class_data_dir", required=True)
    ap.add_argument("--num_class_images", type=int, default=100)
    ap.add_argument("--sample_batch_size", type=int, default=4)
    ap.add_argument("--num_inference_steps", type=int, default=50)
    ap.add_argument("--resolution", type=int, default=512)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if torch.cuda.is_available() else torch.device("cpu")
    dtype = torch.bfloat16 if device.type == "cuda" else torch.float32
    pipe = StableDiffusionPipeline.from_pretrained(args.model_path).to(device, dtype)
    n = generate(pipe, args.class_prompt, args.class_data_dir, args.num_class_images, args.sample_batch_size, rank,
                 world, num_inference_steps=args.num_inference_steps, height=args.resolution, width=args.resolution)
    print(f"[rank {rank}] wrote {n} class images to {args.class_data_dir}")


if __name__ == "__main__":
    main()
