"""Tensor-parallel linears whose collective runs INSIDE the GEMM kernel (NVLink peer memory).

    column parallel (sequence-parallel input):   y        = all_gather(x_shard) · Wᵀ      → AG→GEMM
    row parallel    (sequence-parallel output):  y_shard  = reduce_scatter(x · Wᵀ)        → GEMM→RS

One kernel does both the math and the transfer (``csrc/gemm_sm100.cu`` COMM_AG / COMM_RS): copy
CTAs push the local row blocks to the peers while tcgen05 tiles of the local rows are already running; the RS epilogue
stores partial tiles straight into the owner's staging slot over NVLink and the owners reduce as
arrivals are counted.  The backward passes are the duals (AG→GEMM ↔ GEMM→RS) plus a wgrad that
re-gathers the sharded operand with a P2P push kernel.  The NCCL versions in
``libai_b200/parallel/mappings.py`` are the oracle and the baseline
(reference: libai/layers/linear.py:123-149 does GEMM→all-reduce / identity→all-reduce).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from libai_b200.ops import count_launch, load_ext
from libai_b200.ops.functional import _ACT_IDS
from libai_b200.parallel.symm_mem import CommWorkspace, get_workspace

_N_COMM_CTAS = 16
_BM = 128


class _TPState:
    """Shape-keyed symmetric buffers + counters of one TP group."""

    def __init__(self, group):
        self.ws: CommWorkspace = get_workspace(group)
        self.world, self.rank = self.ws.world, self.ws.rank
        self.ag: Dict[Tuple[int, int], dict] = {}
        self.rs: Dict[Tuple[int, int], dict] = {}

    # gathered operand buffers [2 parities][M, K] + per-row-block arrival counters in the symmetric flag buffer
    # (the pushing peer bumps them)
    def ag_state(self, M: int, K: int) -> dict:
        key = (M, K)
        if key not in self.ag:
            buf = self.ws.buffer(("ag", M, K), 2 * M * K * 2)
            offs = [self.ws.alloc_flags(M // _BM), self.ws.alloc_flags(M // _BM)]
            self.ag[key] = dict(buf=buf, calls=0, flag_offs=offs, targets=[0, 0])
        return self.ag[key]

    # staging [2 parities][world][M/world, N] + arrival counters living in the symmetric flag buffer
    def rs_state(self, M: int, N: int) -> dict:
        key = (M, N)
        if key not in self.rs:
            buf = self.ws.buffer(("rs", M, N), 2 * M * N * 2)
            mbpr = M // self.world // _BM
            offs = [self.ws.alloc_flags(mbpr), self.ws.alloc_flags(mbpr)]
            self.rs[key] = dict(buf=buf, calls=0, flag_offs=offs, targets=[0, 0])
        return self.rs[key]


_STATES: Dict[int, _TPState] = {}


def _state(group) -> _TPState:
    if id(group) not in _STATES:
        _STATES[id(group)] = _TPState(group)
    return _STATES[id(group)]


def fused_supported(M: int, N: int, K: int, world: int) -> bool:
    return M % (_BM * world) == 0 and N % 256 == 0 and K % 8 == 0 and world <= 8


# --------------------------------------------------------------------------------------------------
# raw fused ops
# --------------------------------------------------------------------------------------------------
def ag_gemm(x_shard: torch.Tensor, w: torch.Tensor, bias, act, group, layout: int = 0):
    """``act(all_gather(x_shard) @ op(w) + bias)``; returns ``(y [M, N], x_full [M, K] view)``.
    ``layout`` 0: w is [N, K]; 1: w is [K, N]."""
    ext = load_ext()
    st = _state(group)
    rows, K = x_shard.shape
    M = rows * st.world
    s = st.ag_state(M, K)
    par = s["calls"] & 1
    s["calls"] += 1
    gathered = s["buf"].view(torch.bfloat16, (2, M, K))[par]
    gathered[st.rank * rows : (st.rank + 1) * rows].copy_(x_shard)
    epoch = st.ws.next_epoch()
    s["targets"][par] += 1          # one arrival per remote row block and call (a whole block is copied by one CTA)
    y = ext.gemm_comm(
        gathered, w, layout, bias, _ACT_IDS[act], 1, st.world, st.rank, epoch, s["targets"][par],
        s["buf"].peer_ptrs(par * M * K * 2), st.ws.flags.peer_ptrs(s["flag_offs"][par]), None, None, 0, _N_COMM_CTAS,
    )
    count_launch()
    return y, gathered


def gemm_rs(x: torch.Tensor, w: torch.Tensor, bias, residual, group, layout: int = 0):
    """``reduce_scatter(x @ op(w)) (+ bias) (+ residual)`` → ``[M / world, N]``."""
    ext = load_ext()
    st = _state(group)
    M = x.shape[0]
    N = w.shape[0] if layout == 0 else w.shape[1]
    s = st.rs_state(M, N)
    par = s["calls"] & 1
    s["calls"] += 1
    s["targets"][par] += st.world * (N // 64)
    epoch = st.ws.next_epoch()
    y = ext.gemm_comm(
        x, w, layout, bias, 0, 2, st.world, st.rank, epoch, s["targets"][par],
        s["buf"].peer_ptrs(0), st.ws.flags.peer_ptrs(s["flag_offs"][par]), None, residual, par * M * N, 0,
    )
    count_launch()
    return y


def p2p_all_gather(x_shard: torch.Tensor, group) -> torch.Tensor:
    """All-gather along dim 0 by P2P stores into every peer's symmetric buffer."""
    ext = load_ext()
    st = _state(group)
    rows = x_shard.shape[0]
    tail = tuple(x_shard.shape[1:])
    nbytes = x_shard.numel() * x_shard.element_size()
    key = ("p2pag", nbytes, str(x_shard.dtype))
    s = st.ag.setdefault(key, dict(buf=st.ws.buffer(key, 2 * nbytes * st.world), calls=0))
    par = s["calls"] & 1
    s["calls"] += 1
    epoch = st.ws.next_epoch()
    ext.p2p_allgather(x_shard.contiguous(), s["buf"].peer_ptrs(par * nbytes * st.world), st.ws.flags.peer_ptrs(0),
                      st.ws.done_counter[1:2], st.world, st.rank, epoch)
    count_launch()
    return s["buf"].view(x_shard.dtype, (2, rows * st.world) + tail)[par]


# --------------------------------------------------------------------------------------------------
# autograd wrappers
# --------------------------------------------------------------------------------------------------
class ColumnParallelLinearFused(torch.autograd.Function):
    """fwd AG→GEMM (+bias, +act); bwd dx_shard = GEMM→RS(dy, W), dW = dyᵀ · all_gather(x_shard)."""

    @staticmethod
    def forward(ctx, x_shard, w, bias, act, group):
        ext = load_ext()
        if act not in (None, "none"):
            # keep the pre-activation for backward: run the GEMM without activation, apply it separately
            pre, _ = ag_gemm(x_shard, w, bias, None, group)
            y = ext.bias_act_fwd(pre, None, _ACT_IDS[act])
            count_launch()
            ctx.save_for_backward(x_shard, w, pre)
        else:
            y, _ = ag_gemm(x_shard, w, bias, None, group)
            ctx.save_for_backward(x_shard, w, None)
        ctx.act, ctx.group, ctx.has_bias = act, group, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        ext = load_ext()
        x_shard, w, pre = ctx.saved_tensors
        gy = gy.contiguous()
        if ctx.act not in (None, "none"):
            gy = ext.act_bwd(gy, pre, _ACT_IDS[ctx.act])
            count_launch()
        gx = gw = gb = None
        x_full = p2p_all_gather(x_shard, ctx.group) if ctx.needs_input_grad[1] else None
        if ctx.needs_input_grad[0]:
            gx = gemm_rs(gy, w, None, None, ctx.group, layout=1)  # dy [M, N_loc] · W [N_loc, K]
        if ctx.needs_input_grad[1]:
            main_grad = getattr(w, "main_grad", None)
            if main_grad is not None:
                ext.gemm(gy, x_full, 2, None, main_grad, True, torch.float32)
            else:
                gw = ext.gemm(gy, x_full, 2, None, None, False, torch.float32).to(w.dtype)
            count_launch()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = ext.colsum(gy)
            count_launch()
        return gx, gw, gb, None, None


class RowParallelLinearFused(torch.autograd.Function):
    """fwd y_shard = GEMM→RS(x, W) (+bias +residual in the reduce phase);
    bwd dx = AG→GEMM(dy_shard, W) and dW = all_gather(dy_shard)ᵀ · x (the gathered dy comes for free)."""

    @staticmethod
    def forward(ctx, x, w, bias, residual, group):
        y = gemm_rs(x, w, bias, residual, group)
        ctx.save_for_backward(x, w)
        ctx.group, ctx.has_bias, ctx.has_res = group, bias is not None, residual is not None
        return y

    @staticmethod
    def backward(ctx, gy_shard):
        ext = load_ext()
        x, w = ctx.saved_tensors
        gy_shard = gy_shard.contiguous()
        gx = gw = gb = None
        gx, gy_full = ag_gemm(gy_shard, w, None, None, ctx.group, layout=1)  # dy_full [M, N] · W [N, K_loc]
        if ctx.needs_input_grad[1]:
            main_grad = getattr(w, "main_grad", None)
            if main_grad is not None:
                ext.gemm(gy_full, x, 2, None, main_grad, True, torch.float32)
            else:
                gw = ext.gemm(gy_full, x, 2, None, None, False, torch.float32).to(w.dtype)
            count_launch()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = ext.colsum(gy_shard)  # partial over token shards: summed over TP by sync_gradients
            count_launch()
        return (gx if ctx.needs_input_grad[0] else None), gw, gb, (gy_shard if ctx.has_res else None), None


def column_parallel_linear(x_shard, w, bias, act, group):
    return ColumnParallelLinearFused.apply(x_shard, w, bias, act, group)


def row_parallel_linear(x, w, bias, residual, group):
    return RowParallelLinearFused.apply(x, w, bias, residual, group)
