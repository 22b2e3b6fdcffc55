"""CPU, world_size 2, gloo: the ensemble sharding and the one collective of the rollout (ace_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_members, out_q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ace_b200 import parallel

    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    sl = parallel.member_slice(n_members, r, w)
    # every member's "diagnostic" is a function of its GLOBAL index: the gather must restore global order
    members = torch.arange(n_members, dtype=torch.float32)[sl]
    local = torch.stack([members, members * members], dim=1)  # [B_local, 2]
    full = parallel.gather_members(local)
    t = parallel.max_over_ranks(1.0 + rank)
    parallel.barrier()
    out_q.put((rank, full.tolist(), t, (sl.start, sl.stop)))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_shard_and_gather():
    world, n_members = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_members, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    expect = [[float(i), float(i * i)] for i in range(n_members)]
    assert [r[3] for r in results] == [(0, 3), (3, 6)]
    for rank, full, t, _ in results:
        assert full == expect, (rank, full)
        assert t == 2.0  # max over ranks of 1 + rank


def test_member_slice_rejects_remainder_and_single_process_identity():
    from ace_b200 import parallel

    with pytest.raises(ValueError):
        parallel.member_slice(5, 0, 2)
    assert parallel.member_slice(8, 3, 4) == slice(6, 8)
    x = torch.randn(3, 4)
    assert parallel.gather_members(x) is x
    assert parallel.max_over_ranks(3.5) == 3.5
