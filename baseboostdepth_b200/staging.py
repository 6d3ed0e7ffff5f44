"""Host -> device staging of a training batch as ONE DMA transfer, double-buffered.

The reference moves a batch with ~20 separate ``tensor.to(device)`` calls (``trainer.py:287-289``).
Here every tensor of the batch lives at a fixed offset of one pinned host arena, so the whole batch
crosses PCIe as a single ``cudaMemcpyAsync`` on a side stream into one of two device arenas while
the previous batch is still being consumed; the compute stream only waits on an event.  A
``DataLoader`` collate function can write straight into ``stager.host`` views.

Host-side hazard (and how it is closed): a non-blocking DMA keeps *reading* its pinned source until it
completes, so the host must not rewrite an arena whose upload is still in flight.  There is one pinned
arena per device slot; ``stager.host`` hands out the views of the arena the *next* ``upload_async`` will
send and first waits (``wait_host_free``) until the DMA that last read that arena has finished.  With two
slots the host can therefore fill batch i+2 while batch i+1 is still crossing PCIe.
"""
from __future__ import annotations

import ctypes
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
    """``template``: key -> tensor.  float32 entries cross PCIe as they are.  uint8 entries are frames
    as decoded: they cross at one byte per colour sample and are expanded on the device to the fp32
    tensor ``torchvision.transforms.ToTensor`` would have produced on the host (``x / 255``,
    bit-identical; reference ``datasets/mono_dataset.py:55,201-203``) by one ``bbd_u8_to_f32`` launch
    on the copy stream, so the expansion also overlaps the previous step."""

    def __init__(self, template: Dict, device, slots: int = 2, backend=None):
        self.device = torch.device(device)
        self.keys, self.meta = [], {}
        f_off = u_off = 0
        for k, v in template.items():
            if not torch.is_tensor(v):
                continue
            assert v.dtype in (torch.float32, torch.uint8), "fp32 tensors or 8-bit frames"
            n = v.numel()
            self.keys.append(k)
            if v.dtype == torch.uint8:
                self.meta[k] = ("u8", u_off, n, tuple(v.shape))
                u_off += (n + 255) // 256 * 256             # 256-byte aligned sub-buffers
            else:
                self.meta[k] = ("f32", f_off, n, tuple(v.shape))
                f_off += (n * 4 + 255) // 256 * 256
        self.f32_bytes, self.u8_bytes = f_off, u_off
        self.total = f_off + u_off                          # bytes: [ fp32 entries | 8-bit entries ]
        self.host_arenas = [torch.empty(self.total, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.dev_arena = [torch.empty(self.total, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        # fp32 expansion of the 8-bit part of each slot
        self.expanded = [torch.empty(max(u_off, 1), dtype=torch.float32, device=self.device) for _ in range(slots)]
        self._host_views = [{k: self._view(a, None, k) for k in self.keys} for a in self.host_arenas]
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self._in_flight = [False] * slots                   # an upload out of host arena s was started and not awaited
        self.consumed = [torch.cuda.Event() for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slot = -1
        self._backend = backend
        for views in self._host_views:                      # every arena starts out holding the template batch
            for k, v in template.items():
                if torch.is_tensor(v):
                    views[k].copy_(v.detach().cpu())

    @property
    def next_slot(self):
        return (self.slot + 1) % len(self.dev_arena)

    def wait_host_free(self, slot=None):
        """Block the host until the DMA that last read host arena ``slot`` (default: the one the next
        ``upload_async`` sends) has completed, i.e. until that arena may be rewritten."""
        s = self.next_slot if slot is None else slot
        if self._in_flight[s]:
            self.ready[s].synchronize()
            self._in_flight[s] = False

    @property
    def host(self):
        """Pinned host views (key -> tensor) of the batch the next ``upload_async`` will send.  Safe to
        write: waits first for the previous upload out of this arena."""
        self.wait_host_free()
        return self._host_views[self.next_slot]

    def _view(self, arena, expanded, k):
        kind, off, n, shape = self.meta[k]
        if kind == "f32":
            return arena[off:off + 4 * n].view(torch.float32).view(shape)
        if expanded is None:                                # host side: the bytes themselves
            return arena[self.f32_bytes + off:self.f32_bytes + off + n].view(shape)
        return expanded[off:off + n].view(shape)

    @property
    def nbytes(self):
        return self.total

    def upload_async(self):
        """Start copying the host arena into the next device slot (side stream); returns the slot."""
        self.slot = (self.slot + 1) % len(self.dev_arena)
        s = self.slot
        self.copy_stream.wait_event(self.consumed[s])       # the previous user of this slot is done
        with torch.cuda.stream(self.copy_stream):
            self.dev_arena[s].copy_(self.host_arenas[s], non_blocking=True)
            if self.u8_bytes:
                from . import _lib
                be = self._backend if self._backend is not None else _lib.cuda_backend()
                be.call("u8_to_f32", ctypes.c_void_p(self.dev_arena[s].data_ptr() + self.f32_bytes),
                        ctypes.c_void_p(self.expanded[s].data_ptr()), ctypes.c_size_t(self.u8_bytes))
            self.ready[s].record(self.copy_stream)
        self._in_flight[s] = True
        return s

    def views(self, slot):
        """Device tensors (all fp32) of a slot; the current stream waits until its upload has landed."""
        torch.cuda.current_stream().wait_event(self.ready[slot])
        return {k: self._view(self.dev_arena[slot], self.expanded[slot], k) for k in self.keys}

    def release(self, slot):
        """Call after the last kernel reading the slot was enqueued."""
        self.consumed[slot].record(torch.cuda.current_stream())
