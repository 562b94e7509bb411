"""Run an exported T5 graph once (reference libai/onnx_export/onnx_inference/t5_onnx_infer.py)."""
import argparse

import numpy as np
import torch

from libai_b200.onnx_export.t5_to_onnx import example_inputs

class OnnxModel:
    """onnxruntime session over an exported graph; ``forward(list_of_numpy_inputs)`` feeds the graph inputs in order
    (reference libai/onnx_export/onnx_inference/t5_onnx_infer.py:24-55)."""

    def __init__(self, onnx_filename, providers=None, ort_optimize: bool = True):
        import onnxruntime as ort

        opt = ort.SessionOptions()
        opt.graph_optimization_level = (ort.GraphOptimizationLevel.ORT_ENABLE_EXTENDED if ort_optimize
                                        else ort.GraphOptimizationLevel.ORT_DISABLE_ALL)
        if providers is None:
            providers = [p for p in ("TensorrtExecutionProvider", "CUDAExecutionProvider", "CPUExecutionProvider")
                         if p in ort.get_available_providers()]
        self.sess = ort.InferenceSession(onnx_filename, sess_options=opt, providers=providers)

    def forward(self, input_list):
        feeds = {spec.name: np.asarray(arr) for spec, arr in zip(self.sess.get_inputs(), input_list)}
        return self.sess.run([], feeds)

    __call__ = forward


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True)
    args = ap.parse_args()
    inputs = example_inputs()
    if args.model.endswith(".onnx"):
        import onnxruntime as ort

        sess = ort.InferenceSession(args.model, providers=["CUDAExecutionProvider", "CPUExecutionProvider"])
        out = sess.run(None, {k: v.numpy() for k, v in inputs.items()})[0]
    else:
        out = torch.jit.load(args.model)(*inputs.values()).detach().float().numpy()
    print("logits", out.shape, float(np.abs(out).mean()))
