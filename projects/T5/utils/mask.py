"""Mask helpers (reference projects/T5/utils/mask.py)."""
import torch


def extended_mask(mask, is_decoder=False):
    """``[b, s]`` / ``[b, q, k]`` → boolean ``[b, 1, q, k]`` (lower-triangular when ``is_decoder``)."""
    m = mask.bool()
    if m.dim() == 2:
        m = m[:, None, :] & m[:, :, None]
    m = m[:, None]
    if is_decoder:
        m = m & torch.ones(m.shape[-2:], dtype=torch.bool, device=m.device).tril()
    return m


from projects.MT5.layers.mask_layer import ExtendedMask  # noqa: E402,F401  (module form of ``extended_mask``)
