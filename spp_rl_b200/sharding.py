"""Host-side plumbing for running a population over several GPUs (one process per GPU).

Agents are independent (the reference runs them as separate processes, train/spp_sac_hopper.py:115), so a
population shards over ranks with NO data-path collective; torch.distributed is used only to agree on timing
(max over ranks) and to collect small loss tables.  Works with the gloo backend on CPU (tests) and nccl on GPUs.
"""
import torch


def agent_shard(total_agents: int, rank: int, world: int) -> range:
    """Contiguous block of agent ids owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(total_agents, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value: float, dist=None, device="cpu") -> float:
    """The slowest rank's time: what every multi-GPU throughput number is divided by."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_tables(local: torch.Tensor, total_agents: int, dist=None) -> torch.Tensor:
    """Concatenate per-agent tables ([agents_local, ...]) from all ranks in agent order (rank 0..world-1)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [len(agent_shard(total_agents, r, world)) for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
