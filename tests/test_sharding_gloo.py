"""CPU, world_size 2 over gloo: the multi-GPU host logic (agent sharding, max-over-ranks timing, table gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spp_rl_b200.sharding import agent_shard, gather_tables, max_over_ranks


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = agent_shard(total, rank, world)
    local = torch.tensor([[float(a), float(a) * 2] for a in shard])
    table = gather_tables(local, total, dist)
    slow = max_over_ranks(1.0 + rank, dist)
    q.put((rank, list(shard), table.tolist(), slow))
    dist.destroy_process_group()


def test_agent_shards_cover_population_exactly():
    for total in (1, 7, 148, 256, 1024):
        for world in (1, 2, 3, 8):
            ids = [a for r in range(world) for a in agent_shard(total, r, world)]
            assert ids == list(range(total))
            sizes = [len(agent_shard(total, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_and_timing():
    world, total, port = 2, 7, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[float(a), float(a) * 2] for a in range(total)]
    for rank, shard, table, slow in res:
        assert table == expect            # every rank sees the whole population's table in agent order
        assert slow == 2.0                # max over ranks, not the local time
    assert res[0][1] + res[1][1] == list(range(total))
