"""Multi-GPU data parallelism for the forward path: one process per GPU, the batch sharded by
image, weights replicated, ONE collective per batch — an all-gather of the logits over NCCL
(NVLink 5 / NVSwitch).  Images are independent in eval mode (BatchNorm uses moving statistics,
tasks/image_classification.py:21), so nothing else is exchanged.

The reference has no distributed code at all (its only "collective" is the stub
``all_gather(data): return [data]`` at tasks/human_pose_estimation.py:373-374).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(device_type: str | None = None):
    """Initialise ``torch.distributed`` from the torchrun environment (RANK, LOCAL_RANK, WORLD_SIZE,
    MASTER_ADDR, MASTER_PORT).  Returns ``(rank, local_rank, world_size)``; a single process needs no
    process group."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        use_cuda = (device_type or ("cuda" if torch.cuda.is_available() else "cpu")) == "cuda"
        if use_cuda:
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("gloo")
    return rank, local_rank, world


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous image range ``[lo, hi)`` of rank ``rank``: sizes differ by at most one, order preserved."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local: torch.Tensor, n_total: int | None = None, group=None) -> torch.Tensor:
    """All-gather per-rank row blocks (logits ``(n_r, classes)`` or predictions ``(n_r,)``) into the
    global batch order on every rank.  Equal shards use one ``all_gather_into_tensor``; ragged shards
    are padded to the largest shard and trimmed."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if n_total is None:
        counts = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        sizes = [int(c.item()) for c in all_counts]
    else:
        sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    biggest = max(sizes)
    padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * biggest: r * biggest + sizes[r]] for r in range(world)], dim=0)


class ShardedForward:
    """Run ``module`` on this rank's shard of a global batch and gather the rows.

    ``module`` is any B200-backed ``tlxcv_b200.nn.Module`` returning ``(N, classes)`` logits or
    ``(N,)`` predictions.  ``forward(global_batch)`` accepts the full batch (host or device) and slices
    this rank's images; ``forward_local(shard)`` takes the shard directly (the bench's path)."""

    def __init__(self, module, group=None):
        self.module, self.group = module, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def forward_local(self, shard, n_total=None):
        return gather_rows(self.module(shard), n_total, self.group)

    def forward(self, global_batch):
        n = global_batch.shape[0]
        lo, hi = shard_bounds(n, self.rank, self.world)
        shard = global_batch[lo:hi]
        if shard.device.type != "cuda":
            shard = shard.to("cuda", non_blocking=True)
        return self.forward_local(shard.contiguous(), n)

    __call__ = forward


def pin_rank_affinity(local_rank: int, local_world: int) -> list[int]:
    """Give each rank of a box its own contiguous slice of the host cores this process may use (the H2D staging and the
    launch thread of eight ranks otherwise migrate over the same cores).  Returns the cores now allowed."""
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:      # not Linux
        return []
    if local_world <= 1 or len(cores) < local_world:
        return cores
    per = len(cores) // local_world
    mine = cores[local_rank * per:(local_rank + 1) * per]
    os.sched_setaffinity(0, mine)
    return mine


class OverlappedGather:
    """Forward of step i on the compute stream while the logits of step i-1 are all-gathered on a side stream.

    The all-gather of one step's logits (1 MB per rank) is pure latency (~0.1 ms over NVLink at 8 GPUs); issued on the
    compute stream after every forward it costs 3 % of the step.  Here every step writes its logits into one of ``depth``
    slots, records an event, and the collective of that slot runs on ``gather_stream`` behind the event while the compute
    stream is already in the next forward.  ``step`` only waits (on the device) for the gather that last used the slot
    it is about to overwrite."""

    def __init__(self, plan, world: int, device, depth: int = 2, group=None):
        import torch.cuda as cuda

        self.plan, self.world, self.group, self.depth = plan, world, group, depth
        self.local = [plan.alloc_outputs() for _ in range(depth)]
        rows = self.local[0][0].shape[0]
        self.gathered = [torch.empty((world * rows,) + tuple(self.local[0][0].shape[1:]), dtype=self.local[0][0].dtype,
                                     device=device) for _ in range(depth)] if world > 1 else None
        self.gather_stream = cuda.Stream(device) if world > 1 else None
        self.ev_fwd = [cuda.Event() for _ in range(depth)]
        self.ev_gathered = [cuda.Event() for _ in range(depth)]
        self._i = 0

    def step(self, inputs):
        """Enqueue one forward (+ the gather of its logits); returns the slot index."""
        slot = self._i % self.depth
        self._i += 1
        cur = torch.cuda.current_stream()
        if self.world > 1:
            cur.wait_event(self.ev_gathered[slot])        # the gather that read this slot's logits last time has finished
        self.plan.run(inputs, self.local[slot], graph=True)
        if self.world > 1:
            self.ev_fwd[slot].record(cur)
            with torch.cuda.stream(self.gather_stream):
                self.gather_stream.wait_event(self.ev_fwd[slot])
                dist.all_gather_into_tensor(self.gathered[slot], self.local[slot][0], group=self.group)
                self.ev_gathered[slot].record(self.gather_stream)
        return slot

    def result(self, slot):
        """Gathered rows of ``slot`` (device tensor); the caller's stream waits for the collective."""
        if self.world == 1:
            return self.local[slot][0]
        torch.cuda.current_stream().wait_event(self.ev_gathered[slot])
        return self.gathered[slot]

    def drain(self):
        """Make the current stream wait for every outstanding gather (end of a timed region)."""
        if self.world > 1:
            for ev in self.ev_gathered:
                torch.cuda.current_stream().wait_event(ev)
