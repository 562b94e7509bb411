"""Text classification pipeline (reference libai/inference/text_classification.py:24-140)."""
import torch

from libai_b200.inference.basic import BasePipeline


def _scores_to_records(logits, id2label, num_labels, function_to_apply, return_all_scores):
    if function_to_apply is not None:
        function_to_apply = function_to_apply.lower()
        assert function_to_apply in ("sigmoid", "softmax", "none"), (
            f"Unrecognized `function_to_apply` argument: {function_to_apply}"
        )
    else:
        function_to_apply = "sigmoid" if num_labels == 1 else "softmax"
    if function_to_apply == "sigmoid":
        scores = torch.sigmoid(logits)
    elif function_to_apply == "softmax":
        scores = torch.softmax(logits, dim=-1)
    else:
        scores = logits
    scores = scores.detach().float().cpu().numpy()
    if return_all_scores:
        return [{"label": id2label[i], "score": score.item()} for i, score in enumerate(scores)]
    return {"label": id2label[int(scores.argmax())], "score": scores.max().item()}


class TextClassificationPipeline(BasePipeline):
    def update_cfg(self, data_parallel=1, tensor_parallel=1, pipeline_parallel=1, pipeline_stage_id=None,
                   pipeline_num_layers=None):
        super().update_cfg(data_parallel, tensor_parallel, pipeline_parallel, pipeline_stage_id, pipeline_num_layers)
        mcfg = self.cfg.model.cfg
        mcfg.hidden_dropout_prob = 0.0
        mcfg.attention_probs_dropout_prob = 0.0
        assert "num_labels" in mcfg, "The model's config must contain num_labels"
        if "label2id" not in mcfg:
            label2id = {"Label_" + str(i): i for i in range(mcfg.num_labels)}
            mcfg["label2id"] = label2id
            mcfg["id2label"] = {ind: label for label, ind in label2id.items()}

    def _parse_parameters(self, **pipeline_parameters):
        return {}, {}, {**pipeline_parameters}

    def preprocess(self, inputs, pad: bool = False, **kwargs) -> dict:
        input_ids = torch.tensor(self.tokenizer.encode(inputs), dtype=torch.long).unsqueeze(0)
        return {"input_ids": self.to_device(input_ids),
                "attention_mask": self.to_device(torch.ones_like(input_ids, dtype=torch.bool))}

    def forward(self, model_input_dict) -> dict:
        return self.model(**model_input_dict)

    def postprocess(self, model_outputs_dict, function_to_apply=None, return_all_scores=False, **kwargs) -> dict:
        mcfg = self.cfg.model.cfg
        key = "logits" if "logits" in model_outputs_dict else "prediction_scores"
        id2label = {int(k): v for k, v in dict(mcfg.id2label).items()}
        return _scores_to_records(model_outputs_dict[key][0], id2label, mcfg.num_labels, function_to_apply,
                                  return_all_scores)
