"""Batched greedy validation of a trained policy (reference: trainer/validate.py:60-297, one
instance at a time, batch 1, env on the CPU and both actors on the GPU; driven 100 times per evaluation by
Run.py:672-814 and test_all.py:260-334).

Here the S instances are one env batch on the device and every rollout step is one forward of each actor over all of
them.  The reference's modules are never put in eval mode, so a batch-1 validation normalises every BatchNorm with the
statistics of that ONE instance's rows; `per_instance_batchnorm=True` keeps exactly that by giving each
instance its own BatchNorm group, which is what makes the results comparable with the reference's (and with the
shipped result CSV rows `PPO-G` / `new12800`).  With False the statistics run over the whole batch (the training-rollout
convention at large env batches).
"""
from __future__ import annotations

import numpy as np
import torch

from .env import MASK_ESA, BatchedMTFJSPEnv


def load_actor_state_dicts(npz_path_or_dict, tag):
    """(operation-actor state_dict, machine-actor state_dict) stored as `<tag>/op/<key>` / `<tag>/mch/<key>` arrays
    (tests/golden/policy_golden.npz holds the reference's shipped J6M6E2 checkpoints in that form)."""
    g = np.load(npz_path_or_dict) if isinstance(npz_path_or_dict, str) else npz_path_or_dict
    pick = lambda part: {k.split("/", 2)[2]: torch.as_tensor(g[k]) for k in g.files if k.startswith("%s/%s/" % (tag, part))}
    return pick("op"), pick("mch")


def greedy_validate(job_actor, machine_actor, instances, weights=(0.4, 0.4, 0.2), left_shift=True,
                    per_instance_batchnorm=True, chunk=None, return_actions=False):
    """Greedy rollout of every instance of `instances` ({"t","p","transT","edge"} arrays, [S,...]) to completion.

    -> dict: final4 [S,4] f64 = (makespan, processing energy / N, transport time, idle time) read after `done`
             (validate.py:273-277), objective [S] = w_mk * mk + w_ec * (pt + idle) + w_tt * tt (validate.py:283),
             cumulative rewards [S,5] (validate.py:248-252) and optionally the action sequence [N,S,2].
    Reward weights are the fixed evaluation weights (`reset(Random_weight_type="eval")`, validate.py:142); job mask =
    the ESA rule (`Eval_esa_update_...`, algorithm/ppo_algorithm.py:321-417); both actors take the arg-max
    (agent_func.py:41-51, 65-73; ties resolve to the lowest index as torch.max / torch.argmax do)."""
    t = np.asarray(instances["t"])
    S, N, M = t.shape
    J = N // M
    E = np.asarray(instances["edge"]).shape[1]
    tf32_job = getattr(job_actor, "precision", "fp32") == "tf32"
    if chunk is None:
        chunk = 1 if (tf32_job and per_instance_batchnorm) else S   # the tcgen05 encoder has one BatchNorm group per call
    final4 = np.zeros((S, 4))
    cum = np.zeros((S, 5))
    acts = np.zeros((N, S, 2), dtype=np.int32) if return_actions else None
    env = None
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        B = s1 - s0
        if env is None or env.B != B:
            env = BatchedMTFJSPEnv(B, J, M, E, left_shift=left_shift, weights=weights, obs_dtype=torch.float32,
                                   mask_mode=MASK_ESA)
        env.load(*(np.asarray(instances[k])[s0:s1] for k in ("t", "p", "transT", "edge")))
        env.scaler_init()
        env.reset(np.tile(np.asarray(weights, dtype=np.float64), (B, 1)))
        env.obs(MASK_ESA)
        groups = B if per_instance_batchnorm else 1
        h_mch = None
        tot = torch.zeros((B, 5), dtype=torch.float64, device=env.device)
        with torch.no_grad():
            for step in range(N):
                if groups > 1 and tf32_job:
                    raise ValueError("per-instance BatchNorm on the tcgen05 job actor needs chunk=1")
                prob, pooled, _ = job_actor.evaluate(env.task_fea, env.adj_w, env.adj_src, env.candidate, h_mch,
                                                     env.job_mask, groups=groups)
                a = prob.argmax(dim=-1)
                env.op.copy_(env.candidate.long().gather(1, a.unsqueeze(-1)).squeeze(-1).to(torch.int32))
                m1, mmask = env.mfea1(env.op)
                mp, h_mch, _ = machine_actor.forward(m1, env.mach_fea, pooled, mmask, groups=groups)
                env.mach.copy_(mp.argmax(dim=-1).to(torch.int32))
                if acts is not None:
                    acts[step, s0:s1, 0] = env.op.cpu().numpy(); acts[step, s0:s1, 1] = env.mach.cpu().numpy()
                env.step_obs(env.op, env.mach, MASK_ESA)
                tot += env.reward5
        assert bool(env.done.all()) and not bool(env.invalid.any())
        final4[s0:s1] = env.costs().cpu().numpy()
        cum[s0:s1] = tot.cpu().numpy()
    w = weights
    out = {"final4": final4, "objective": w[0] * final4[:, 0] + w[1] * (final4[:, 1] + final4[:, 3]) + w[2] * final4[:, 2],
           "cumulative_rewards": cum}
    if acts is not None:
        out["actions"] = acts
    return out
