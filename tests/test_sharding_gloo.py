"""CPU, world_size 2 over gloo: the multi-GPU host logic (agent sharding, max-over-ranks timing, table gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import numpy as np

from spp_rl_b200.sharding import (agent_shard, allreduce_adv_stats, env_shard_rows, epoch_local_minibatches_device, gather_tables, local_minibatch,
                                  local_minibatch_device, max_over_ranks)


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


def test_env_sharding_covers_rows_and_minibatches():
    """SPP-PPO data parallelism: every global row has exactly one owner, local order is step-major with stride E/W, and the
    ranks' slices of a global minibatch partition it."""
    E, T = 12, 5
    for world in (1, 2, 3, 4):
        owned = [env_shard_rows(E, T, r, world) for r in range(world)]
        assert sorted(np.concatenate(owned).tolist()) == list(range(E * T))
        El = E // world
        for r, rows in enumerate(owned):
            assert rows.reshape(T, El)[:, 0].tolist() == [t * E + r * El for t in range(T)]       # env r*El at every step
            # local row ids of a permutation slice point back at the same global rows
            perm = np.random.RandomState(world + r).permutation(E * T)[:23]
            loc = local_minibatch(perm, E, r, world)
            mine = [g for g in perm.tolist() if (g % E) // El == r]
            assert rows[loc].tolist() == mine
            assert local_minibatch_device(torch.from_numpy(perm), E, r, world).tolist() == loc.tolist()      # the torch form (device path)
        full = np.random.RandomState(9).permutation(E * T)
        for r in range(world):      # a whole epoch filtered at once == the minibatches filtered one by one (ragged last minibatch)
            ids, off = epoch_local_minibatches_device(torch.from_numpy(full), 23, E, r, world)
            for k, b0 in enumerate(range(0, E * T, 23)):
                assert ids[off[k]:off[k + 1]].tolist() == local_minibatch(full[b0:b0 + 23], E, r, world).tolist()
        sizes = [len(local_minibatch(np.arange(E * T), E, r, world)) for r in range(world)]
        assert sum(sizes) == E * T


def _adv_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = np.random.RandomState(0).randn(1000)[rank::world]            # this rank's advantages
    g = allreduce_adv_stats([a.size, a.sum(), (a * a).sum()], dist)
    q.put((rank, g.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_advantage_statistics():
    """the fp64 (n, sum, sum of squares) all-reduce reproduces torch.std (unbiased) of the undivided advantage vector"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_adv_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = np.random.RandomState(0).randn(1000)
    for _, (n, s1, s2) in res:
        mean = s1 / n
        std = np.sqrt((s2 - n * mean * mean) / (n - 1))
        assert n == 1000 and abs(mean - a.mean()) < 1e-12
        assert abs(std - float(torch.from_numpy(a).std())) < 1e-12
