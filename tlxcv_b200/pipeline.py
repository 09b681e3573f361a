"""Host-buffer streaming through one plan: the end-to-end call a user of the reference's API makes.

The reference's ``model(x)`` takes a host (or device) NCHW fp32 tensor and returns logits.  At
>50 k img/s the 154 MB/step host->device copy of a bs256 fp32 batch is as long as the forward
itself, so the public streaming API pipelines three stages on three streams:

    copy stream     H2D of batch i+1          (pinned host -> device input slot)
    compute stream  forward of batch i        (CUDA-graph replay of the plan)
    output stream   gather + D2H of batch i-1 (NCCL all-gather of the logits when sharded, then the
                                               device -> pinned host copy of what THIS rank hands back)

Events order the stages per slot; nothing is cached across steps: every step moves its own inputs
H2D from pinned memory and its own result D2H.  (Round 1 ran the gather and the D2H on the compute
stream, and every rank copied the whole gathered block to its host: at 8 GPUs that serialised 0.1 ms
of NCCL latency and 8 x 8 MB of D2H per step into the forward.)
"""
from __future__ import annotations

import torch

from . import runtime


class HostPipeline:
    def __init__(self, module, batch_shape, depth: int = 2, device=None, gather=None, dtype=torch.float32,
                 gather_to: str = "all"):
        """``batch_shape``: (N, C, H, W) of every submitted fp32 batch, or (N, H, W, C) with ``dtype=torch.uint8`` for a
        module that starts with ``vision.NormalizeToTensor``.  ``gather``: optional callable applied to the device
        result on the output stream (e.g. ``tlxcv_b200.dist.gather_rows``).  ``gather_to``: which ranks copy the
        GATHERED rows to their host buffer — ``"all"`` or ``"rank0"`` (the other ranks then copy only their own rows
        into the leading part of ``host_out``)."""
        self.device = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.depth = depth
        self.gather = gather
        if gather_to not in ("all", "rank0"):
            raise ValueError("gather_to must be 'all' or 'rank0'")
        self.gather_to = gather_to
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.dev_in = [torch.empty(batch_shape, dtype=dtype, device=self.device) for _ in range(depth)]
        self.plan, self.structure, _ = runtime.get_plan(module, (self.dev_in[0],), {})
        if self.plan.n_in != 1 or self.plan.n_out != 1:
            raise NotImplementedError("HostPipeline handles single-input single-output modules")
        self.dev_out = [self.plan.alloc_outputs() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.out_stream = torch.cuda.Stream(self.device)
        self.ev_h2d = [torch.cuda.Event() for _ in range(depth)]       # input slot filled
        self.ev_fwd = [torch.cuda.Event() for _ in range(depth)]       # forward done: input slot free, result ready
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]       # result slot drained (gathered + copied out)
        self._i = 0
        self.h2d_bytes = self.dev_in[0].numel() * self.dev_in[0].element_size()
        self.d2h_bytes = 0      # bytes this rank copied to its host in the last submit

    def submit(self, host_in: torch.Tensor, host_out: torch.Tensor):
        """Enqueue one batch: pinned ``host_in`` -> device -> forward -> (gather) -> pinned ``host_out``.

        With ``gather`` set, ``host_out`` holds the gathered rows on the ranks ``gather_to`` names and this rank's own
        rows (leading part of the buffer) on the others."""
        if not host_in.is_pinned() or not host_out.is_pinned():
            raise runtime.B200RuntimeError("HostPipeline needs pinned host tensors")
        slot = self._i % self.depth
        self._i += 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_fwd[slot])        # the slot's previous forward has consumed its input
            self.dev_in[slot].copy_(host_in, non_blocking=True)
            self.ev_h2d[slot].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.ev_h2d[slot])
            self.compute_stream.wait_event(self.ev_out[slot])     # the slot's previous result has left the device buffer
            out = self.plan.run([self.dev_in[slot]], self.dev_out[slot], graph=True)[0]
            self.ev_fwd[slot].record(self.compute_stream)
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(self.ev_fwd[slot])
            if self.gather is not None and (self.gather_to == "all" or self.rank == 0):
                src = self.gather(out)
                dst = host_out
            elif self.gather is not None:
                self.gather(out)                                   # collective: every rank takes part
                src, dst = out, host_out[: out.shape[0]]
            else:
                src, dst = out, host_out
            dst.copy_(src, non_blocking=True)
            self.d2h_bytes = src.numel() * src.element_size()
            self.ev_out[slot].record(self.out_stream)

    def synchronize(self):
        self.copy_stream.synchronize()
        self.compute_stream.synchronize()
        self.out_stream.synchronize()
