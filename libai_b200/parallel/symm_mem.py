"""Symmetric (peer-mapped) device memory: the substrate of the in-kernel NVLink collectives.

Each rank of a process group allocates the same number of bytes with ``cudaMalloc`` (outside the
caching allocator), the 64-byte CUDA IPC handles are exchanged through ``torch.distributed``, and
every rank maps all peers' buffers.  Kernels receive the table of peer pointers and use plain
``ld/st/red`` on them – the traffic goes over NVLink 5 / NVSwitch.  
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from libai_b200.ops import load_ext


class SymmetricBuffer:
    """``nbytes`` of zero-initialised device memory mapped by every rank of ``group``."""

    def __init__(self, nbytes: int, group=None, tag: str = ""):
        ext = load_ext()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nbytes = int((nbytes + 255) // 256 * 256)
        self.local = ext.symm_alloc(self.nbytes)  # uint8 tensor owning the allocation
        self.tag = tag
        if self.world == 1:
            self.ptrs = [self.local.data_ptr()]
            return
        handle = ext.symm_export(self.local)
        payload = bytes(handle.numpy().tobytes())
        gathered: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(gathered, payload, group=group)
        self.ptrs = []
        for r, h in enumerate(gathered):
            if r == self.rank:
                self.ptrs.append(self.local.data_ptr())
            else:
                ht = torch.frombuffer(bytearray(h), dtype=torch.uint8)
                self.ptrs.append(int(ext.symm_open(ht)))
        dist.barrier(group=group)

    def close(self) -> None:
        """Unmap the peers' allocations and free the local one.  Collective over ``group``: nobody frees memory a
        peer may still have mapped (a later ``cudaMalloc`` can hand the same range out again, and re-opening its IPC
        handle while the stale mapping exists fails with "resource already mapped")."""
        if self.local is None:
            return
        ext = load_ext()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)           # every rank's kernels on these buffers have retired
            for r, p in enumerate(self.ptrs):
                if r != self.rank:
                    ext.symm_close(int(p))
            dist.barrier(group=self.group)           # every mapping is gone before any owner frees
        self.ptrs = []
        self.local = None

    def view(self, dtype: torch.dtype, shape, offset_bytes: int = 0) -> torch.Tensor:
        """Typed view of the *local* buffer."""
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty((), dtype=dtype).element_size()
        return self.local[offset_bytes : offset_bytes + nb].view(dtype).view(*shape)

    def peer_ptrs(self, offset_bytes: int = 0) -> List[int]:
        return [p + offset_bytes for p in self.ptrs]


class CommWorkspace:
    """Per-group bookkeeping for the fused collectives: handshake flags, arrival counters, epochs and
    shape-keyed data buffers (all symmetric)."""

    FLAG_BYTES = 1 << 16

    def __init__(self, group):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.flags = SymmetricBuffer(self.FLAG_BYTES, group, "flags")
        self._flag_cursor = 64 * 4 * self.world  # first slots are reserved for the handshake rows
        self.epoch = 0
        self.device = self.flags.local.device
        self.done_counter = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._bufs: Dict[Tuple, object] = {}

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def alloc_flags(self, n_words: int) -> int:
        """Reserve ``n_words`` uint32 counters in the symmetric flag buffer; returns the byte offset
        (identical on every rank because all ranks allocate in the same order)."""
        off = self._flag_cursor
        self._flag_cursor += (n_words * 4 + 63) // 64 * 64
        assert self._flag_cursor <= self.FLAG_BYTES, "symmetric flag buffer exhausted"
        return off

    def buffer(self, key: Tuple, nbytes: int) -> SymmetricBuffer:
        if key not in self._bufs:
            self._bufs[key] = SymmetricBuffer(nbytes, self.group, str(key))
        return self._bufs[key]

    def close(self) -> None:
        """Release every buffer of this workspace (collective; same order on every rank of the group)."""
        for key in list(self._bufs):
            buf = self._bufs.pop(key)
            if isinstance(buf, SymmetricBuffer):
                buf.close()
        self.flags.close()


_WORKSPACES: Dict[int, CommWorkspace] = {}


def get_workspace(group) -> CommWorkspace:
    key = id(group)
    if key not in _WORKSPACES:
        _WORKSPACES[key] = CommWorkspace(group)
    return _WORKSPACES[key]


def release_workspaces() -> None:
    """Tear down every workspace (before the process groups they belong to are destroyed or re-made).  Must be called
    by all ranks at the same point; objects still holding views of the buffers must be dropped first."""
    for key in list(_WORKSPACES):
        _WORKSPACES.pop(key).close()
