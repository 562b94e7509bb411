"""Greedy text generation from an exported GPT-2 graph (reference libai/onnx_export/onnx_inference/
gpt2_onnx_infer.py): runs the ``.onnx`` file with onnxruntime, or the TorchScript ``.pt`` fallback."""
import argparse

import numpy as np
import torch


class OnnxModel:
    """onnxruntime session over an exported graph; ``forward(list_of_numpy_inputs)`` feeds the graph inputs in order
    (reference libai/onnx_export/onnx_inference/gpt2_onnx_infer.py:24-55)."""

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


class ExportedLM:
    def __init__(self, path):
        self.path = path
        if path.endswith(".onnx"):
            import onnxruntime as ort

            self.sess = ort.InferenceSession(path, providers=["CUDAExecutionProvider", "CPUExecutionProvider"])
            self.run = lambda ids: self.sess.run(None, {"input_ids": ids.astype(np.int64)})[0]
        else:
            self.module = torch.jit.load(path)
            self.run = lambda ids: self.module(torch.from_numpy(ids.astype(np.int64))).detach().float().numpy()

    def generate(self, ids, max_new_tokens=16, eos_token_id=None):
        ids = np.asarray(ids, dtype=np.int64)[None]
        for _ in range(max_new_tokens):
            nxt = int(self.run(ids)[0, -1].argmax())
            ids = np.concatenate([ids, [[nxt]]], axis=1)
            if eos_token_id is not None and nxt == eos_token_id:
                break
        return ids[0].tolist()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True)
    ap.add_argument("--vocab-file", required=True)
    ap.add_argument("--merges-file", required=True)
    ap.add_argument("--prompt", default="a dog")
    args = ap.parse_args()
    from libai_b200.tokenizer import GPT2Tokenizer

    tok = GPT2Tokenizer(args.vocab_file, args.merges_file)
    out = ExportedLM(args.model).generate(tok.encode(args.prompt), eos_token_id=tok.eos_token_id)
    print(tok.decode(out))
