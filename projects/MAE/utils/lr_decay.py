"""Parameter groups for MAE (reference projects/MAE/utils/lr_decay.py): layer-wise learning-rate decay for
fine-tuning (``lr_scale = decay^(L+1-layer)``) and the plain decay / no-decay split for pre-training."""
import json


def get_layer_idx_for_vit(name, num_layers):
    if name in ("cls_token", "pos_embed") or name.startswith("patch_embed"):
        return 0
    if name.startswith("blocks"):
        return int(name.split(".")[1]) + 1
    return num_layers


def param_groups_lrd(model, weight_decay=0.05, layer_decay=0.75, no_weight_decay_list=None):
    skip = set(no_weight_decay_list if no_weight_decay_list is not None else getattr(model, "no_weight_decay", lambda: ())())
    num_layers = len(model.blocks) + 1
    scales = [layer_decay ** (num_layers - i) for i in range(num_layers + 1)]
    groups = {}
    for n, p in model.named_parameters():
        if not p.requires_grad or p.device.type == "meta":
            continue
        decay = 0.0 if (p.ndim == 1 or n in skip) else weight_decay
        layer = get_layer_idx_for_vit(n, num_layers)
        key = f"layer_{layer}_{'no_decay' if decay == 0.0 else 'decay'}"
        groups.setdefault(key, {"lr_scale": scales[layer], "weight_decay": decay, "params": []})["params"].append(p)
    return list(groups.values())


def param_groups_weight_decay(model, weight_decay=1e-5, skip_list=()):
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad or p.device.type == "meta":
            continue
        (no_decay if (p.ndim <= 1 or n.endswith(".bias") or n in skip_list) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]
