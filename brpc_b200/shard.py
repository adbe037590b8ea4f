"""Multi-GPU sharding of the hot path.

The path has no exchange step: all parser state hangs off the Socket, so connections
are partitioned `gpu = SlotOfVRefId(socket_id) % n_gpus` (the low 32 bits of a SocketId are
the resource-pool slot, src/brpc/versioned_ref_with_id.h:55-66) and every GPU runs the whole
pipeline on its own connections.  The only collective is the bvar-style reduction of the
per-GPU counters (bvar::Reducer::get_value -> AgentCombiner::combine_agents,
src/bvar/reducer.h:227-233, detail/combiner.h:243-253): sum for Adders, max/min for
Maxer/Miner — one all-reduce of an int64[8], NCCL on GPUs, any torch.distributed backend in tests.
"""
import numpy as np


def owner_of(socket_ids, world_size):
    """Rank that serves each connection."""
    return (np.asarray(socket_ids, dtype=np.uint64) & np.uint64(0xffffffff)) % np.uint64(max(1, world_size))


def my_runs(runs, rank, world_size):
    """Indices of the runs this rank owns (order preserved: per-socket order is what matters)."""
    return np.nonzero(owner_of(runs["socket_id"], world_size) == rank)[0]


def reduce_counters(local_counters, op="sum", device=None):
    """All-reduce an int64[k] counter vector over the default process group.
    op: "sum" (bvar::Adder), "max" (Maxer), "min" (Miner)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(local_counters, dtype=np.int64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
    return t.cpu().numpy()


def max_over_ranks(values, device=None):
    """Device-side timing rule: a multi-GPU number is the max over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()
