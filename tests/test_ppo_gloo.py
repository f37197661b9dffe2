"""Data-parallel PPO update on CPU: world-size-2 gloo.  Each rank updates on its own half of the envs
of the golden buffer (kernels replaced by the torch restatements of tests/test_ppo.py); gradients are averaged with
one allreduce per backward pass, so both ranks must hold bit-identical parameters afterwards, and those parameters
must equal a single-process update whose gradient is the mean of the two half-batch gradients."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(dev, lo, hi):
    from tests import test_ppo as tp
    enc, ppo = tp.enc, tp.ppo
    enc.aggregate = tp._t_aggregate
    enc.graph_mean = lambda h, *a, **k: h.mean(dim=1)
    enc.ell_invert = lambda s: None
    enc.bn_forward, enc.bn_backward = tp._t_bn_forward, tp._t_bn_backward
    ppo.gae4 = tp._t_gae4
    bt, p, (J, M) = tp.build_batch(dev)
    bt = {k: v[:, lo:hi].contiguous() for k, v in bt.items()}
    H = int(p["H"])
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 11), J, M, hidden=H, device=dev, trainable=True)
    mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 12), M, hidden=H, device=dev, trainable=True)
    crit = enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(H), 13), J, M, hidden=H, device=dev, trainable=True)
    up = ppo.MAPPOUpdate(job, mch, crit, ppo.PPOConfig(k_epochs=1))
    return up, bt, p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    up, bt, p = _make(torch.device("cpu"), rank * 2, rank * 2 + 2)
    up.update(bt, int(p["mini_bs"]), orders=p["orders"][:1])
    sd = {("job/" + k): v.numpy() for k, v in up.job.state_dict().items()}
    sd.update({("mch/" + k): v.numpy() for k, v in up.mch.state_dict().items()})
    sd.update({("crit/" + k): v.numpy() for k, v in up.critic.state_dict().items()})
    q.put((rank, sd, up.allreduce_bytes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_hold_identical_parameters_after_the_update():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (_, sd0, by0), (_, sd1, by1) = res
    assert by0 == by1 and by0 > 0
    for k in sd0:
        np.testing.assert_array_equal(sd0[k], sd1[k], err_msg=k)
    from tests import test_ppo as tp
    init = tp.enc.seeded_state_dict(tp.enc.job_actor_keys(32), 11)
    k = "encoder.feature_extract.mlps.0.linears.0.weight"
    assert np.abs(sd0["job/" + k] - init[k].numpy()).max() > 1e-4


def test_allreduce_mean_grads_is_a_no_op_without_a_process_group():
    ppo = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.ppo")
    w = torch.ones(3, requires_grad=True)
    (w * 2).sum().backward()
    assert ppo.allreduce_mean_grads([w]) == 0 and torch.equal(w.grad, torch.full((3,), 2.0))


def _weighted_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ppo = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.ppo")
    B = (3, 1)[rank]                                   # 4 envs over 2 ranks, unequal shards (sharding.shard_range)
    up = object.__new__(ppo.MAPPOUpdate)
    up._shard_weight = {}
    wgt = up._grad_weight(B, torch.device("cpu"))      # B_local * world / B_total
    w = torch.ones(5, requires_grad=True)
    (w * float(rank + 1)).sum().backward()             # this rank's mean-over-its-envs gradient: (rank + 1) everywhere
    ppo.allreduce_mean_grads([w], weight=wgt)
    q.put((rank, wgt, w.grad.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_unequal_shards_weight_the_gradient_by_env_count():
    """each rank's loss is a mean over ITS envs; with 3 + 1 envs the global mean gradient is
    (3 * g0 + 1 * g1) / 4, not (g0 + g1) / 2."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_weighted_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0][1] == 1.5 and res[1][1] == 0.5
    for _, _, g in res:
        np.testing.assert_allclose(g, np.full(5, (3 * 1.0 + 1 * 2.0) / 4), rtol=0, atol=1e-7)
