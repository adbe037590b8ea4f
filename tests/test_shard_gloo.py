"""CPU test of the N>1 path with world_size 2 over gloo: connections are sharded by socket id,
each rank runs the path on its shard (here: the oracle stands in for the GPU), and the bvar-style
counters all-reduce to what a single process sees on the whole set."""
import os
import random
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import _oracle as O
    from _traffic import SEED, mixed_frames, split_runs
    from brpc_b200 import make_runs, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(SEED)                      # same traffic on every rank
    streams = [mixed_frames(rng, rng.randrange(1, 30)) for _ in range(37)]
    data, runs = make_runs(split_runs(rng, streams))
    runs["socket_id"] = np.arange(len(runs), dtype=np.uint64) * 3 + (7 << 32)     # version bits above the slot
    mine = shard.my_runs(runs, rank, world)
    rs, msgs, resp = O.process_batch(O.make_config(), data, runs[mine])
    local = np.array([rs["consumed"].sum(), rs["n_msgs"].sum(), msgs["resp_len"].sum(), len(mine),
                      (rs["parse_error"] != 2).sum(), 1, 0, 0], dtype=np.int64)
    total = shard.reduce_counters(local, "sum")
    biggest = shard.reduce_counters(np.array([rs["n_msgs"].max() if len(rs) else 0]), "max")
    t = shard.max_over_ranks([1.0 + rank])
    if rank == 0:
        frs, fmsgs, fresp = O.process_batch(O.make_config(), data, runs)
        expect = np.array([frs["consumed"].sum(), frs["n_msgs"].sum(), fmsgs["resp_len"].sum(), len(runs),
                           (frs["parse_error"] != 2).sum(), world, 0, 0], dtype=np.int64)
        q.put((total.tolist(), expect.tolist(), int(biggest[0]), int(frs["n_msgs"].max()), float(t[0])))
    dist.destroy_process_group()


def test_sharded_counters_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    total, expect, biggest, biggest_expect, tmax = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60); assert p.exitcode == 0
    assert total == expect
    assert biggest == biggest_expect and tmax == 2.0


def test_owner_of_uses_slot_bits():
    sys.path.insert(0, ROOT)
    from brpc_b200 import shard
    ids = np.array([(5 << 32) | 10, (9 << 32) | 10, 11, (1 << 32) | 12], dtype=np.uint64)
    assert shard.owner_of(ids, 4).tolist() == [2, 2, 3, 0]
    assert shard.owner_of(ids, 1).tolist() == [0, 0, 0, 0]
