"""Host -> device staging of a training batch as ONE DMA transfer, double-buffered.

The reference moves a batch with ~20 separate ``tensor.to(device)`` calls (``trainer.py:287-289``).
Here every tensor of the batch lives at a fixed offset of one pinned host arena, so the whole batch
crosses PCIe as a single ``cudaMemcpyAsync`` on a side stream into one of two device arenas while
the previous batch is still being consumed; the compute stream only waits on an event.  A
``DataLoader`` collate function can write straight into ``stager.host`` views.
"""
from __future__ import annotations

import os
from typing import Dict

import torch


def bind_to_gpu_numa(device_index: int) -> str:
    """Pin this process to the CPU cores next to the GPU (NVML affinity) *before* allocating pinned
    memory, so the staging arena is NUMA-local to the GPU's PCIe root.  Returns a short description;
    does nothing (and says so) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [i * 64 + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "no overlap between NVML affinity and the allowed CPUs"
        os.sched_setaffinity(0, allowed)
        return f"bound to {len(allowed)} cores next to GPU {device_index}"
    except Exception as exc:  # noqa: BLE001
        return f"not bound ({type(exc).__name__})"


class BatchStager:
    def __init__(self, template: Dict, device, slots: int = 2):
        self.device = torch.device(device)
        self.keys, self.meta, off = [], {}, 0
        for k, v in template.items():
            if not torch.is_tensor(v):
                continue
            assert v.dtype == torch.float32, "the loss path is fp32"
            n = v.numel()
            self.keys.append(k)
            self.meta[k] = (off, n, tuple(v.shape))
            off += (n + 63) // 64 * 64                      # 256-byte aligned sub-buffers
        self.total = off
        self.host_arena = torch.empty(self.total, dtype=torch.float32).pin_memory()
        self.host = {k: self._view(self.host_arena, k) for k in self.keys}
        self.dev_arena = [torch.empty(self.total, dtype=torch.float32, device=self.device) for _ in range(slots)]
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [torch.cuda.Event() for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slot = -1
        for k, v in template.items():
            if torch.is_tensor(v):
                self.host[k].copy_(v.detach().cpu())

    def _view(self, arena, k):
        off, n, shape = self.meta[k]
        return arena[off:off + n].view(shape)

    @property
    def nbytes(self):
        return self.total * 4

    def upload_async(self):
        """Start copying the host arena into the next device slot (side stream); returns the slot."""
        self.slot = (self.slot + 1) % len(self.dev_arena)
        s = self.slot
        self.copy_stream.wait_event(self.consumed[s])       # the previous user of this slot is done
        with torch.cuda.stream(self.copy_stream):
            self.dev_arena[s].copy_(self.host_arena, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        return s

    def views(self, slot):
        """Device tensors of a slot; the current stream waits until its upload has landed."""
        torch.cuda.current_stream().wait_event(self.ready[slot])
        return {k: self._view(self.dev_arena[slot], k) for k in self.keys}

    def release(self, slot):
        """Call after the last kernel reading the slot was enqueued."""
        self.consumed[slot].record(torch.cuda.current_stream())
