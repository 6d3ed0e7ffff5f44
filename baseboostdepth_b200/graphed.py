"""The loss step as a replayed CUDA graph, one graph per staging slot.

``loss_step`` + ``backward`` is a fixed sequence of ~11 kernels of ours and a dozen tiny tensor ops on
buffers whose addresses do not change from step to step once the batch lives in a
``staging.BatchStager`` slot.  ``GraphedLossStep`` captures that sequence once per slot and replays it:
the host cost of a step drops from ~1 ms of Python and launches to one ``cudaGraphLaunch``, and the
loss of step *i* is read back while step *i+1* is already running (an asynchronous copy into pinned
memory + an event), so neither the upload, nor the launch overhead, nor the read-back sits on the
critical path.

    stager = BatchStager(template, device)
    step = GraphedLossStep(stager, make_io, opt, plan)
    slot = stager.upload_async()
    for ...:
        fill(stager.host)                    # write the next batch (waits until that pinned arena is free)
        nxt = stager.upload_async()          # ... which crosses PCIe during this step
        step.launch(slot)                    # replay; gradients land in step.grads(slot)
        loss_of_previous = step.collect()    # host float of the step launched before this one (or None)
        slot = nxt
    last = step.collect(final=True)

``make_io(views) -> (inputs, outputs, leaves)`` maps the stager's device views of one slot to the
trainer's ``inputs`` / ``outputs`` dictionaries and names the tensors that need gradients.
``prepare_inputs(inputs)`` (optional) does the same for ``inputs`` (e.g. the colour pyramid levels derived
from the uploaded full-resolution frame instead of being uploaded).
``prepare(outputs)`` (optional) runs inside the captured step on a copy of ``outputs``: tensor ops that
derive further entries from the uploaded ones, like the error-induced poses of ``--decomp``
(``trainer.py:376-377``).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .trainer import loss_step


class GraphedLossStep:
    def __init__(self, stager, make_io: Callable, opt, plan, num_scales: Optional[int] = None, warmup: int = 3,
                 prepare: Optional[Callable] = None, prepare_inputs: Optional[Callable] = None):
        self.stager = stager
        self._graphs, self._loss, self._leaves = [], [], []
        n_slots = len(stager.dev_arena)
        dev = stager.device
        self._host_loss = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(n_slots)]
        self._done = [torch.cuda.Event() for _ in range(n_slots)]
        self._pending = []                                   # slots launched and not yet collected, oldest first
        for slot in range(n_slots):
            s = stager.upload_async()
            assert s == slot
            views = stager.views(slot)                       # waits for the upload on the current stream
            inputs, outputs, leaves = make_io(views)

            def run():
                for p in leaves.values():
                    p.grad = None
                outs, ins = dict(outputs), dict(inputs)
                if prepare is not None:                      # e.g. the decomp error poses, derived from T in-graph
                    prepare(outs)
                if prepare_inputs is not None:               # e.g. the colour pyramid, derived from the uploaded frame
                    prepare_inputs(ins)
                losses = loss_step(ins, outs, opt, plan, noise=None, num_scales=num_scales)
                losses["loss"].backward()
                return losses["loss"]

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):                      # caches tables / weights, warms the allocator
                    run()
            torch.cuda.current_stream().wait_stream(side)
            for p in leaves.values():
                p.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = run()
            stager.release(slot)
            self._graphs.append(g)
            self._loss.append(loss)
            self._leaves.append(leaves)
        torch.cuda.synchronize(dev)

    def launch(self, slot: int) -> None:
        """Replay the step on the batch in ``slot`` (its upload is awaited on the current stream)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self.stager.ready[slot])
        self._graphs[slot].replay()
        self._host_loss[slot].copy_(self._loss[slot].detach(), non_blocking=True)
        self._done[slot].record(cur)
        self.stager.release(slot)
        self._pending.append(slot)

    def collect(self, final: bool = False):
        """Loss (host float) of the oldest launched step that is not the one just launched; with
        ``final`` the remaining ones are drained and the last value is returned."""
        value = None
        while self._pending and (final or len(self._pending) > 1):
            slot = self._pending.pop(0)
            self._done[slot].synchronize()
            value = float(self._host_loss[slot])
        return value

    def grads(self, slot: int) -> Dict:
        """Gradient tensors (static buffers of the slot's graph) keyed like ``leaves``."""
        return {k: p.grad for k, p in self._leaves[slot].items()}
