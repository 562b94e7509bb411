"""Text → image with DALL-E 2 under data/tensor parallelism.

Spec: reference projects/DALLE2/dalle2_inference.py:33-184 — a ``BasePipeline`` that tokenises the captions, samples
image embeddings from the prior, decodes 64×64 images, optionally saves them and upsamples 4×/16× with SwinIR.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 projects/DALLE2/dalle2_inference.py \
        --tensor_parallel 4 --save_images --output_dir ./outputs --upsample_scale 4
"""
import argparse
import os
import sys
from typing import Dict

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from libai_b200.config import instantiate  # noqa: E402
from libai_b200.inference.basic import BasePipeline  # noqa: E402
from libai_b200.utils import distributed as dist  # noqa: E402


class Dalle2Pipeline(BasePipeline):
    def load_pretrain_weight(self, libai_cfg_model, model_path, mode=None):
        from projects.DALLE2.dalle2.dalle2_loader import Dalle2ModelLoader

        model = instantiate(libai_cfg_model)
        return Dalle2ModelLoader(model, libai_cfg_model, model_path).load()

    def _place(self, model):
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device.startswith("cuda") else torch.device("cpu")
        return model.to(dev)

    def build_tokenizer(self, cfg):
        from projects.DALLE2.dalle2.tokenizer import SimpleTokenizer

        try:
            return SimpleTokenizer()
        except FileNotFoundError as e:          # the CLIP BPE vocabulary file is not shipped (no network here)
            import logging

            logging.getLogger(__name__).warning("CLIP BPE vocabulary missing (%s): set `pipeline.tokenizer` yourself", e)
            return None

    _FORWARD_DEFAULTS = dict(num_samples_per_batch=2, prior_cond_scale=1.0, decoder_cond_scale=3.5)
    _POST_KEYS = ("save_images", "upsample_scale", "output_dir", "swinir_path")

    def _parse_parameters(self, **kwargs):
        """Only what the caller passed: call-time values override the constructor's, nothing else does."""
        forward = {k: kwargs[k] for k in self._FORWARD_DEFAULTS if k in kwargs}
        post = {k: kwargs[k] for k in self._POST_KEYS if k in kwargs}
        return {}, forward, post

    def split_data(self, text):
        """Captions are split over the data-parallel ranks (balanced, contiguous)."""
        topo = dist.get_dist_util()
        n, world, rank = len(text), topo.data_parallel_size, topo.dp_rank
        lo, hi = n * rank // world, n * (rank + 1) // world
        return text[lo:hi]

    def preprocess(self, input_, **kwargs) -> dict:
        text = self.split_data(list(input_))
        return {"text": text, "tokens": self.to_device(self.tokenizer.tokenize(text))}

    def forward(self, model_input_dict, **fp) -> dict:
        tokens = model_input_dict["tokens"]
        m = self.model
        _, text_encodings, text_mask = m.prior.clip.embed_text(tokens)
        fp = {**self._FORWARD_DEFAULTS, **fp}
        image_embed = m.prior.sample(tokens, num_samples_per_batch=fp["num_samples_per_batch"], cond_scale=fp["prior_cond_scale"])
        images = m.decoder.sample(image_embed=image_embed, text_encodings=text_encodings, text_mask=text_mask,
                                  cond_scale=fp["decoder_cond_scale"])
        return {"image_embed": images}

    def postprocess(self, model_output_dict, **pp) -> dict:
        if not pp.get("save_images", False):
            return model_output_dict
        from torchvision.transforms.functional import to_pil_image

        out = pp.get("output_dir") or "./outputs"
        os.makedirs(out, exist_ok=True)
        rank = dist.get_dist_util().dp_rank
        images = model_output_dict["image_embed"].float().cpu().clamp(0, 1)
        for i, img in enumerate(images):
            to_pil_image(img).save(f"{out}/r{rank}_{i}.png")
        scale = pp.get("upsample_scale")
        if scale:
            from projects.DALLE2.swinir import load_model, upsample4x, upsample16x

            swinir = load_model(pp.get("swinir_path")).to(self.model.prior.noise_scheduler.betas.device)
            up = (upsample4x if scale == 4 else upsample16x)(images, swinir).cpu()
            for i, img in enumerate(up):
                to_pil_image(img).save(f"{out}/r{rank}_{i}_{scale}x.png")
        print(f"Images have been saved under {out}.")
        return model_output_dict


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--config_file", default="projects/DALLE2/configs/dalle2_config.py")
    p.add_argument("--data_parallel", type=int, default=1)
    p.add_argument("--tensor_parallel", type=int, default=int(os.environ.get("WORLD_SIZE", 1)))
    p.add_argument("--pipeline_parallel", type=int, default=1)
    p.add_argument("--upsample_scale", type=int, choices=[4, 16], default=None,
                   help="upsample scale, if 4x, output resolution will be 256 x 256.")
    p.add_argument("--swinir_path", default="./swinir/weights/003_realSR_BSRGAN_DFOWMFC_s64w8_SwinIR-L_x4_GAN.pth")
    p.add_argument("--output_dir", default="./outputs")
    p.add_argument("--save_images", action="store_true")
    p.add_argument("--mode", default="libai", choices=["libai", "random"])
    return p.parse_args(argv)


if __name__ == "__main__":
    args = parse_args()
    pipe = Dalle2Pipeline(config_file=args.config_file, data_parallel=args.data_parallel,
                          tensor_parallel=args.tensor_parallel, pipeline_parallel=args.pipeline_parallel,
                          mode="random", save_images=args.save_images, upsample_scale=args.upsample_scale,
                          output_dir=args.output_dir, swinir_path=args.swinir_path)
    texts = ["a shiba inu wearing a beret and black turtleneck", "a teddy bear on a skateboard in times square"]
    pipe(texts)
