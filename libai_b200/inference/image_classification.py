"""Image classification pipeline (reference libai/inference/image_classification.py:27-160): the test transform of
``cfg.dataloader.test[0].dataset.transform`` is applied to an image path; labels come from the ImageNet-1k table
when ``num_classes == 1000`` (``Label_i`` otherwise)."""
import os

import torch

from libai_b200.config import instantiate
from libai_b200.inference.basic import BasePipeline
from libai_b200.inference.text_classification import _scores_to_records


class ImageClassificationPipeline(BasePipeline):
    def __init__(self, config_file, data_parallel=None, tensor_parallel=None, pipeline_parallel=None,
                 pipeline_stage_id=None, pipeline_num_layers=None, model_path=None, mode="libai", **kwargs):
        super().__init__(config_file, data_parallel, tensor_parallel, pipeline_parallel, pipeline_stage_id,
                         pipeline_num_layers, model_path, mode, **kwargs)
        if "num_classes" in self.cfg.model:
            self.num_classes = self.cfg.model.num_classes
        elif "cfg" in self.cfg.model and "num_classes" in self.cfg.model.cfg:
            self.num_classes = self.cfg.model.cfg.num_classes
        else:
            raise AttributeError("The model's config must contain num_classes")
        self.id2label = {ind: label for label, ind in self.label2id(self.num_classes).items()}
        self.transform = instantiate(self.cfg.dataloader.test[0].dataset.transform)

    def _parse_parameters(self, **pipeline_parameters):
        return {}, {}, {**pipeline_parameters}

    def preprocess(self, inputs, **kwargs) -> dict:
        from PIL import Image

        assert os.path.exists(inputs), "inputs must be an existing image path!"
        with open(inputs, "rb") as f:
            img = Image.open(f).convert("RGB")
        img = self.transform(img).unsqueeze(0)
        param = next(self.model.parameters())
        return {"images": img.to(device=param.device, dtype=param.dtype)}

    def forward(self, model_input_dict) -> dict:
        return self.model(**model_input_dict)

    def postprocess(self, model_outputs_dict, function_to_apply=None, return_all_scores=False, **kwargs) -> dict:
        return _scores_to_records(model_outputs_dict["prediction_scores"][0], self.id2label, self.num_classes,
                                  function_to_apply, return_all_scores)

    def label2id(self, num_classes):
        """``label → index``: ImageNet-1k names for 1000 classes, generic ``Label_i`` otherwise."""
        from libai_b200.inference.utils.imagenet_class import IMAGENET_LABELS

        if num_classes == 1000:
            return {label: i for i, label in enumerate(IMAGENET_LABELS)}
        return {"Label_" + str(i): i for i in range(num_classes)}
