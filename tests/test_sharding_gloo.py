"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes own
contiguous env-id shards, generate the same action stream a single process would, and reduce the
timing with MAX exactly as bench.py does. No GPU needed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401  (sets sys.path)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    from rogue_gym_python.rollout import env_seeds, shard_range, synthetic_actions
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    ids = np.arange(lo, hi, dtype=np.uint64)
    acts = np.stack([synthetic_actions(t, ids) for t in range(5)])
    seeds = env_seeds(ids)
    # every rank contributes its block; the union must be the single-process stream
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, acts.tobytes(), seeds.tobytes()))
    t = torch.tensor([10.0 + rank, float(hi - lo)], dtype=torch.float64)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((gathered, float(mx[0]), float(sm[1])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from rogue_gym_python.rollout import env_seeds, shard_range, synthetic_actions
    world, n_total = 2, 1001  # odd on purpose: blocks differ by one env
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, t_max, n_sum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == 11.0 and n_sum == n_total          # max-over-ranks time, whole-job env count
    blocks = sorted((lo, hi) for lo, hi, _, _ in gathered)
    assert blocks[0][0] == 0 and blocks[-1][1] == n_total
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))  # contiguous, disjoint, complete
    ids = np.arange(n_total, dtype=np.uint64)
    want = np.stack([synthetic_actions(t, ids) for t in range(5)])
    for lo, hi, acts, seeds in gathered:
        got = np.frombuffer(acts, np.uint8).reshape(5, hi - lo)
        assert np.array_equal(got, want[:, lo:hi])
        assert np.array_equal(np.frombuffer(seeds, np.uint64), env_seeds(ids[lo:hi]))


def test_shard_range_properties():
    from rogue_gym_python.rollout import shard_range
    for n in (0, 1, 7, 65536, 524288, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_action_stream_matches_oracle_side_definition(oracle):
    """The product's generator and the oracle's independent copy of SURVEY §8d agree."""
    from rogue_gym_python.rollout import synthetic_actions
    ids = np.arange(5000, dtype=np.uint64) * 13
    for t in (0, 1, 999, 123456):
        assert np.array_equal(synthetic_actions(t, ids), oracle.synthetic_actions(t, ids))
    assert set(synthetic_actions(3, ids).tobytes()) <= set(b".hjklnbuy>s")
