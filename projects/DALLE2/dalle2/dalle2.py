"""Text → image: prior (text → CLIP image embedding) then decoder (embedding → pixels).
Spec: reference projects/DALLE2/dalle2/models.py:2504-2541."""
import torch
from torch import nn


class DALLE2(nn.Module):
    def __init__(self, *, prior, decoder, prior_num_samples=2, tokenizer=None, prior_weight_path="",
                 decoder_weight_path="", **unused):
        super().__init__()
        self.prior, self.decoder = prior, decoder
        self.prior_num_samples = prior_num_samples
        self.decoder_need_text_cond = decoder.condition_on_text_encodings
        self._tokenizer = tokenizer
        self.prior_weight_path, self.decoder_weight_path = prior_weight_path, decoder_weight_path

    @property
    def tokenizer(self):
        if self._tokenizer is None:
            from .tokenizer import SimpleTokenizer

            self._tokenizer = SimpleTokenizer()
        return self._tokenizer

    @torch.no_grad()
    def forward(self, text, cond_scale=1.0, prior_cond_scale=1.0, return_pil_images=False):
        self.eval()
        device = next(self.parameters()).device
        is_str = isinstance(text, str) or (isinstance(text, (list, tuple)) and all(isinstance(t, str) for t in text))
        one_text = isinstance(text, str) or (not is_str and text.shape[0] == 1)
        if is_str:
            text = self.tokenizer.tokenize([text] if isinstance(text, str) else list(text))
        text = text.to(device)
        image_embed = self.prior.sample(text, num_samples_per_batch=self.prior_num_samples, cond_scale=prior_cond_scale)
        text_encodings = text_mask = None
        if self.decoder_need_text_cond:
            _, text_encodings, text_mask = self.prior.clip.embed_text(text)
        images = self.decoder.sample(image_embed, text_encodings=text_encodings, text_mask=text_mask, cond_scale=cond_scale)
        if return_pil_images:
            from torchvision.transforms.functional import to_pil_image

            images = [to_pil_image(img.float().cpu().clamp(0, 1)) for img in images]
        return images[0] if one_text else images
