"""N > 1 host logic on CPU: world-size-2 gloo.  Each rank builds its env slice; slices must equal the rows of the
unsharded batch, the oracle rollout of a slice (with its env_offset) must equal the same rows of the unsharded
rollout, and the episode-statistics / max-time reductions must agree with the single-process values."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.mtfjsp_oracle import OracleEnv

TOTAL, J, M, E, SEED = 70, 3, 4, 2, 77  # 70 envs over 2 ranks; crosses nothing special but uneven with 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rollout(first, count, d, w):
    o = OracleEnv(count, J, M, E)
    o.load(d["t"], d["p"], d["transT"], d["edge"])
    o.scaler_init()
    o.reset(w)
    out = o.rollout_random(J * M, seed=5, env_offset=first, record_actions=True)
    return out["actions"], o.costs(), o.export_state()["ft"]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.sharding")
    first, count, d, w = sh.make_shard(TOTAL, rank, world, J, M, E, SEED)
    acts, costs, ft = _rollout(first, count, d, w)
    stats = sh.reduce_episode_stats(costs)
    tmax = sh.max_over_ranks(10.0 + rank)
    q.put((rank, first, count, d["t"], w, acts, costs, ft, stats, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slices_equal_the_unsharded_batch():
    sh = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.sharding")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = ins.synthetic_instances(0, TOTAL, J, M, E, SEED)
    w = ins.random_weights(0, TOTAL, SEED)
    acts, costs, ft = _rollout(0, TOTAL, d, w)
    covered = 0
    for rank, first, count, t, wr, a, c, f, stats, tmax in res:
        assert (first, count) == sh.shard_range(TOTAL, rank, world)
        assert first == covered
        covered += count
        np.testing.assert_array_equal(t, d["t"][first:first + count])
        np.testing.assert_array_equal(wr, w[first:first + count])
        np.testing.assert_array_equal(a, acts[:, first:first + count])
        np.testing.assert_array_equal(c, costs[first:first + count])
        np.testing.assert_array_equal(f, ft[first:first + count])
        assert tmax == 11.0
        assert stats["count"] == TOTAL
        np.testing.assert_allclose(stats["mk"], costs[:, 0].mean(), rtol=1e-12)
        obj = 0.4 * costs[:, 0] + 0.4 * (costs[:, 1] + costs[:, 3]) + 0.2 * costs[:, 2]
        np.testing.assert_allclose(stats["objective"], obj.mean(), rtol=1e-12)
    assert covered == TOTAL


def test_shard_ranges_partition_any_batch():
    sh = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.sharding")
    for total in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                first, count = sh.shard_range(total, r, world)
                assert first == nxt
                for e in (first, first + count - 1):
                    if count:
                        assert sh.owner_of(e, total, world) == r
                nxt += count
            assert nxt == total
