"""Run an exported T5 graph once (reference libai/onnx_export/onnx_inference/t5_onnx_infer.py)."""
import argparse

import numpy as np
import torch

from libai_b200.onnx_export.t5_to_onnx import example_inputs

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
