"""Host-buffer streaming through one plan: the end-to-end call a user of the reference's API makes.

The reference's ``model(x)`` takes a host (or device) NCHW fp32 tensor and returns logits.  At
>50 k img/s the 154 MB/step host->device copy of a bs256 fp32 batch is as long as the forward
itself, so the public streaming API double-buffers: while batch i runs, batch i+1 is already
crossing PCIe on a copy stream.  Every step still moves its own inputs H2D from pinned memory and
its own result D2H; nothing is cached across steps.
"""
from __future__ import annotations

import torch

from . import runtime


class HostPipeline:
    def __init__(self, module, batch_shape, depth: int = 2, device=None, gather=None, dtype=torch.float32):
        """``batch_shape``: (N, C, H, W) of every submitted fp32 batch, or (N, H, W, C) with ``dtype=torch.uint8`` for a
        module that starts with ``vision.NormalizeToTensor``.  ``gather``: optional callable applied to the device
        result before the D2H copy (e.g. ``tlxcv_b200.dist.gather_rows``)."""
        self.device = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.depth = depth
        self.gather = gather
        self.dev_in = [torch.empty(batch_shape, dtype=dtype, device=self.device) for _ in range(depth)]
        self.plan, self.structure, _ = runtime.get_plan(module, (self.dev_in[0],), {})
        if self.plan.n_in != 1 or self.plan.n_out != 1:
            raise NotImplementedError("HostPipeline handles single-input single-output modules")
        self.dev_out = [self.plan.alloc_outputs() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.ev_h2d = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self._i = 0
        self.h2d_bytes = self.dev_in[0].numel() * self.dev_in[0].element_size()

    def submit(self, host_in: torch.Tensor, host_out: torch.Tensor):
        """Enqueue one batch: pinned ``host_in`` (N,C,H,W) fp32 -> device -> forward -> pinned ``host_out``."""
        if not host_in.is_pinned() or not host_out.is_pinned():
            raise runtime.B200RuntimeError("HostPipeline needs pinned host tensors")
        slot = self._i % self.depth
        self._i += 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_done[slot])       # the slot's previous forward has consumed its input
            self.dev_in[slot].copy_(host_in, non_blocking=True)
            self.ev_h2d[slot].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.ev_h2d[slot])
            out = self.plan.run([self.dev_in[slot]], self.dev_out[slot], graph=True)[0]
            if self.gather is not None:
                out = self.gather(out)
            host_out.copy_(out, non_blocking=True)
            self.ev_done[slot].record(self.compute_stream)

    def synchronize(self):
        self.copy_stream.synchronize()
        self.compute_stream.synchronize()
