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


# ---- SPP-PPO data parallelism: environments shard over ranks, trajectories never cross GPUs.
# Global rollout rows are step-major, row g = t * E + e; rank r owns environments [r * E/W, (r + 1) * E/W) and keeps them
# step-major with stride E/W, so GAE is local and a global minibatch is the union of the ranks' slices of one permutation.
def env_shard_rows(E: int, T: int, rank: int, world: int):
    """Global row ids of the rows `rank` owns, in its local (step-major, stride E/W) order."""
    import numpy as np

    if E % world:
        raise ValueError("environments (%d) must divide evenly over %d ranks" % (E, world))
    El = E // world
    return (np.arange(T)[:, None] * E + rank * El + np.arange(El)[None, :]).reshape(-1)


def local_minibatch(global_idx, E: int, rank: int, world: int):
    """The part of a global minibatch (global row ids) that `rank` owns, as LOCAL row ids, in minibatch order."""
    import numpy as np

    idx = np.asarray(global_idx, np.int64)
    El = E // world
    e = idx % E
    mine = (e // El) == rank
    return ((idx[mine] // E) * El + e[mine] % El).astype(np.int64)


def local_minibatch_device(global_idx, E: int, rank: int, world: int):
    """local_minibatch for a torch int64 tensor on the device (same result, same order); the boolean selection synchronises
    once to learn the count, which the caller needs on the host anyway."""
    El = E // world
    e = global_idx % E
    mine = torch.div(e, El, rounding_mode="floor") == rank
    return (torch.div(global_idx, E, rounding_mode="floor") * El + e % El)[mine]


def epoch_local_minibatches_device(perm, batch: int, E: int, rank: int, world: int):
    """One epoch at once: `perm` (torch int64, the epoch's global permutation) -> (local row ids of every row this rank owns, in
    permutation order, and the host list of offsets such that minibatch k is ids[off[k]:off[k + 1]]).  A handful of device ops
    and ONE host read per epoch instead of a filter + synchronisation per minibatch."""
    El = E // world
    e = perm % E
    mine = torch.div(e, El, rounding_mode="floor") == rank
    ids = (torch.div(perm, E, rounding_mode="floor") * El + e % El)[mine]
    csum = torch.cumsum(mine.to(torch.int64), 0)
    n = perm.numel()
    ends = torch.arange(batch, n + batch, batch, device=perm.device).clamp_(max=n) - 1
    off = [0] + csum[ends].tolist()
    return ids, off


def allreduce_adv_stats(local_stats, dist=None, device="cpu"):
    """(n, sum, sum of squares) of the local advantages -> the global triple, in fp64 (torch.std parity needs it)."""
    t = torch.as_tensor(local_stats, dtype=torch.float64, device=device).clone()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()


# ---- inside one GPU: update bursts of a population larger than the grid --------------------------------------------------------
BURST_ITEM_STEPS = 2      # kItemSteps of update_burst_interleaved_kernel (csrc/update_kernel.cu)


def burst_items(population: int, grad_steps: int, grid: int, cta: int, item_steps: int = BURST_ITEM_STEPS):
    """The work items CTA `cta` of `grid` takes in an update burst, in order, as (agent, first_step, end_step) -- the host-side
    statement of the loop in update_burst_interleaved_kernel (population > grid) and update_burst_kernel (population <= grid: whole
    agents).  An item may start once `first_step` steps of its agent are complete (the agent's progress word)."""
    if population <= grid:
        return [(a, 0, grad_steps) for a in range(cta, population, grid)]
    chunks = (grad_steps + item_steps - 1) // item_steps
    out = []
    for w in range(cta, population * chunks, grid):
        q, agent = divmod(w, population)
        out.append((agent, q * item_steps, min(grad_steps, (q + 1) * item_steps)))
    return out


def burst_step_times(population: int, grad_steps: int, grid: int, item_steps: int = BURST_ITEM_STEPS) -> int:
    """Length of a burst in units of one update step when every step takes the same time (the longest CTA's item list):
    256 agents x 50 steps on 148 CTAs -> 88 (whole agents per CTA: 100)."""
    return max(sum(e - b for _, b, e in burst_items(population, grad_steps, grid, c, item_steps)) for c in range(min(grid, population)))
