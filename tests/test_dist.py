"""N>1 host logic on the CPU (-m "not gpu"): batch sharding and the row gather over gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tlxcv_b200.dist import gather_rows, shard_bounds


def test_shard_bounds_partition_the_batch():
    for n in (0, 1, 7, 8, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_total * 5, dtype=torch.float32).reshape(n_total, 5)
        lo, hi = shard_bounds(n_total, rank, world)
        got = gather_rows(full[lo:hi].clone(), n_total)
        got2 = gather_rows(full[lo:hi].clone())                     # sizes discovered with a collective
        pred = gather_rows(full[lo:hi, 0].to(torch.int64).clone(), n_total)
        q.put((rank, torch.equal(got, full), torch.equal(got2, full), torch.equal(pred, full[:, 0].to(torch.int64))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])        # equal shards and ragged shards
def test_gather_rows_world2_gloo(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in results)


def test_pin_rank_affinity_gives_disjoint_core_slices():
    from tlxcv_b200.dist import pin_rank_affinity

    if not hasattr(os, "sched_getaffinity"):
        pytest.skip("no sched_getaffinity here")
    saved = os.sched_getaffinity(0)
    try:
        cores = sorted(saved)
        if len(cores) < 2:
            pytest.skip("needs two host cores")
        a = pin_rank_affinity(0, 2)
        os.sched_setaffinity(0, saved)
        b = pin_rank_affinity(1, 2)
        assert a and b and not set(a) & set(b) and set(a) | set(b) <= set(cores)
        os.sched_setaffinity(0, saved)
        assert pin_rank_affinity(0, 1) == cores           # a single rank keeps everything
    finally:
        os.sched_setaffinity(0, saved)
