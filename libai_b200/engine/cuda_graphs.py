"""CUDA-graph capture of the transformer blocks.

A training step of the 24-layer benchmark model issues ≈820 kernels; the host needs ≈25 ms to enqueue them against
≈29 ms of device time, so the step is within ~15 % of being launch-bound (``bench.py`` reports
``host_enqueue_ms_per_step``).  The blocks are where the launches are (≈14 forward + ≈20 backward kernels each), they
have static shapes and no host synchronisation, so each block's forward and backward are captured once into CUDA graphs
(``torch.cuda.make_graphed_callables``) and replayed: 48 graph launches replace ≈800 kernel launches per step.

Why this works with the native backward: the wgrad / bias-grad / LayerNorm-grad kernels accumulate straight into the
optimizer's persistent fp32 ``main_grad`` buffer and return ``None`` to autograd — those kernels (and the TMA
descriptors they were launched with, which are by-value kernel parameters) are part of the captured backward graph and
replay against the same addresses.  Parameters live in the optimizer's flat buffer, i.e. at fixed addresses as well.

Tensor-parallel blocks are captured too when the collectives are the fused ones (``train.dist.fused_tp_comm``): the
AG→GEMM / GEMM→RS kernels keep their call counters, arrival targets and write credits in device memory and advance them
themselves, so a replay is indistinguishable from a fresh launch.

Not captured: embedding, LM head + loss, optimizer, NCCL collectives.  Restrictions (checked): one tensor argument per
block, no pipeline parallelism, no activation checkpointing.  Dropout inside captured blocks draws its Philox offset
from a device counter that the dropout kernels advance (``ops/functional.py``), so replays produce fresh masks.
"""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence

import torch
from torch import nn

from libai_b200 import ops

logger = logging.getLogger(__name__)


def _with_launch_count(block: nn.Module, fwd_kernels: int, bwd_kernels: int) -> nn.Module:
    """The kernels inside a replayed graph are still launches of our kernels: keep ``ops.launch_count()`` honest.
    (``make_graphed_callables`` patches ``block.forward`` in place, so module names / state-dict keys are unchanged.)"""
    graphed_forward = block.forward

    def forward(hidden):
        ops.count_launch(fwd_kernels + (bwd_kernels if torch.is_grad_enabled() and block.training else 0))
        return graphed_forward(hidden)

    block.forward = forward
    return block


def graph_transformer_blocks(blocks: Sequence[nn.Module], sample_hidden: torch.Tensor, warmup_iters: int = 3
                             ) -> Optional[List[nn.Module]]:
    """Capture ``blocks`` (called as ``block(hidden)``) for inputs shaped like ``sample_hidden``.  Returns the graphed
    replacements, or ``None`` (with a log line) when a precondition does not hold."""
    if not sample_hidden.is_cuda:
        return None
    # (dropout inside the blocks is fine: the native kernels take their Philox (seed, offset) from PyTorch's CUDA
    # generator, which hands out graph-safe state during capture and refreshes the offset before every replay)
    # Per-block kernel counts for the launch bookkeeping.  The forward count comes from a no-grad eager call (it also
    # makes sure every lazy one-time setup has happened: cudaFuncSetAttribute, extension load, driver entry points);
    # the backward count from the warm-up + capture passes below.  No eager *backward* here: it would leave
    # AccumulateGrad nodes bound to the current stream, and the capture stream would then depend on uncaptured work
    # (cudaErrorStreamCaptureIsolation).
    with torch.no_grad():
        n0 = ops.launch_count()
        blocks[0](sample_hidden.detach())
        fwd_kernels = ops.launch_count() - n0
    torch.cuda.synchronize()
    samples = tuple((sample_hidden.detach().clone().requires_grad_(True),) for _ in blocks)
    n0 = ops.launch_count()
    graphed = torch.cuda.make_graphed_callables(tuple(blocks), samples, num_warmup_iters=warmup_iters,
                                                allow_unused_input=True)
    per_block = (ops.launch_count() - n0) // ((warmup_iters + 1) * len(blocks))
    ops.count_launch(-(ops.launch_count() - n0))          # warm-up / capture launches are not training launches
    graphed = graphed if isinstance(graphed, (tuple, list)) else (graphed,)
    counts = [(fwd_kernels, max(per_block - fwd_kernels, 0))] * len(blocks)
    return [_with_launch_count(g, f, b) for g, (f, b) in zip(graphed, counts)]


def enable_for_model(model: nn.Module, example_batch: dict) -> bool:
    """Replace the blocks of a ``PipelineStageMixin`` model by graphed versions.  ``example_batch``: one training batch
    (device tensors) used to discover the hidden-state shape.  The gradient buffers must be zeroed afterwards (the
    capture runs backward passes)."""
    from libai_b200.utils import distributed as dutil

    topo = dutil.get_dist_util()
    if (topo.pipeline_parallel_size > 1 or getattr(model, "activation_checkpoint", False)
            or (topo.tensor_parallel_size > 1 and not topo.fused_tp_comm)):
        # (1F1B keeps several micro-batches in flight per block; checkpointing re-runs the forward inside backward;
        # tensor parallelism is captured only in its fused form — AG→GEMM / GEMM→RS kernels whose handshake state lives
        # in device memory (ops/comm_gemm.py) — the NCCL form keeps eager launches)
        logger.warning("cuda graphs: pipeline parallelism / NCCL-form tensor parallelism / activation checkpointing — not captured")
        return False
    layers = model.stage_layers() if hasattr(model, "stage_layers") else None
    if not isinstance(layers, nn.ModuleList) or len(layers) == 0:
        return False
    shapes = {}

    def grab(mod, args, kwargs):
        shapes["hidden"] = args[0]
        shapes["n_args"] = len(args) + len([v for v in kwargs.values() if v is not None])

    h = layers[0].register_forward_pre_hook(grab, with_kwargs=True)
    was_training = model.training
    model.train()
    with torch.no_grad():
        model(**example_batch)
    h.remove()
    if shapes.get("n_args", 0) != 1 or not torch.is_tensor(shapes.get("hidden")):
        logger.warning("cuda graphs: blocks take more than the hidden state — not captured")
        return False
    try:
        new = graph_transformer_blocks(list(layers), shapes["hidden"])
    except Exception as e:   # capture is an optimisation: never take the run down with it
        logger.warning("cuda graphs: capture failed (%s: %s) — staying with eager launches", type(e).__name__, str(e)[:200])
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        model.train(was_training)
        return False
    if new is None:
        model.train(was_training)
        return False
    assert all(a is b for a, b in zip(new, layers)), "make_graphed_callables is expected to patch the modules in place"
    model.train(was_training)
    logger.info("cuda graphs: captured forward+backward of %d blocks (hidden %s)", len(new), tuple(shapes["hidden"].shape))
    return True
