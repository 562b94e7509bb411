"""Initialise a linear-probe ViT from a MoCo pre-training checkpoint (reference projects/MOCOV3/utils/
load_checkpoint.py): keep ``base_encoder.*`` (minus its projector head), drop everything else."""
import torch

from libai_b200.parallel.state import load_full_state_dict


def load_checkpoint(model, path, weight_style="oneflow", num_heads=12):
    obj = torch.load(path if not path.endswith("/") else path + "model", map_location="cpu", weights_only=False)
    sd = obj.get("model", obj.get("state_dict", obj))
    if weight_style == "pytorch":
        from projects.MOCOV3.utils.weight_convert import convert_state_dict

        sd = convert_state_dict({k.replace("module.", ""): v for k, v in sd.items()}, num_heads)
    keep = {}
    for k, v in sd.items():
        if k.startswith("base_encoder.") and not k.startswith("base_encoder.head"):
            keep[k[len("base_encoder."):]] = v
    missing, unexpected, mismatched = load_full_state_dict(model, keep, strict=False)
    assert all(m.startswith("head") for m in missing), missing
    return model
