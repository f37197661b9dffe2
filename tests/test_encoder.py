"""Encoder side: state_dict compatibility with the shipped checkpoints (CPU) and forward parity of
JobActor / MachineActor on the device against golden outputs of the reference modules
(tests/golden/gen_encoder_golden.py), fed from the CUDA environment's native observation layout.
Tolerance: FP32 (rtol 2e-4, atol 2e-5) -- the reference runs these layers in float32 as well."""
import importlib
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_state_dict_layout_matches_shipped_checkpoints():
    enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
    tab = json.load(open(os.path.join(GOLD, "shipped_checkpoint_keys.json")))
    job = {k: list(v) for k, v in enc.job_actor_keys(128).items()}
    mch = {k: list(v) for k, v in enc.machine_actor_keys(128).items()}
    assert job == tab["PPO_operation_actor_J6M6E2_1000.pth"]
    assert mch == tab["PPO_machine_actor_J6M6E2_1000.pth"]


@pytest.mark.gpu
@pytest.mark.parametrize("H", [128, 32])
def test_actor_forward_matches_reference_modules(H):
    torch = pytest.importorskip("torch")
    enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    g = np.load(os.path.join(GOLD, "replay_j6m6_ls_esa.npz"))
    e = np.load(os.path.join(GOLD, "encoder_golden.npz"))
    J, M, E = int(g["J"]), int(g["M"]), int(g["E"])
    B = g["t"].shape[0]
    env = envm.BatchedMTFJSPEnv(B, J, M, E, left_shift=True, obs_dtype=torch.float32)
    env.load(g["t"], g["p"], g["transT"], g["edge"])
    env.scaler_init()
    env.reset(g["weights"][0])
    env.obs(1)
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 11), J, M, hidden=H)
    mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 12), M, hidden=H)
    dev = env.device
    i32 = lambda x: torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32)).to(dev)
    close = lambda a, b, name: np.testing.assert_allclose(a.cpu().numpy(), b, rtol=2e-4, atol=2e-5, err_msg=name)
    cur = -1
    for tag in ("init", "s05", "s20", "s34"):
        k = "H%d_%s_" % (H, tag)
        s = int(e[k + "step"])
        while cur < s:  # replay the dump's actions up to step s
            cur += 1
            act = g["actions"][0, cur]
            env.step_obs(i32(act[:, 0]), i32(act[:, 1]), 1)
        hgm = None if e[k + "hgm_in"].size == 0 else torch.tensor(e[k + "hgm_in"]).to(dev)
        ti, ai, la, prob, pooled, jv = job.forward(env.task_fea, env.adj_w, env.adj_src, env.candidate, hgm,
                                                   env.job_mask, greedy=True)
        close(prob, e[k + "prob"], "job prob " + tag)
        close(pooled, e[k + "pooled"], "job pooled " + tag)
        close(jv, e[k + "job_v"], "job value " + tag)
        np.testing.assert_array_equal(ti.cpu().numpy(), e[k + "task_index"])
        nxt = g["actions"][0, s + 1]
        m1, mmask = env.mfea1(i32(nxt[:, 0]))
        np.testing.assert_array_equal(mmask.cpu().numpy().astype(bool), e[k + "mmask"])
        mp, hp, mv = mch.forward(m1, env.mach_fea, torch.tensor(e[k + "pooled"]).to(dev), mmask)
        close(mp, e[k + "mch_prob"], "machine prob " + tag)
        close(hp, e[k + "mch_pooled"], "machine pooled " + tag)
        close(mv, e[k + "mch_v"], "machine value " + tag)


@pytest.mark.gpu
def test_aggregate_kernel_matches_dense_fp64_reference():
    torch = pytest.importorskip("torch")
    enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    B, J, M, E = 64, 10, 10, 3
    d = ins.synthetic_instances(0, B, J, M, E, 3)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    env.reset(ins.random_weights(0, B, 3))
    for s in range(57):
        env.random_step(seed=1)
    A = env.dense_adj(torch.float64)                       # [B,N,N], adj[dst,src], diagonal 1 (reference layout)
    for C in (12, 128):
        h = torch.randn(B, J * M, C, device=env.device)
        ref = torch.bmm(A, h.double()) / (A != 0).sum(-1, keepdim=True).double()   # gcn_mlp.py:125-149
        out = enc.aggregate(h, env.adj_w, env.adj_src)
        np.testing.assert_allclose(out.cpu().numpy(), ref.float().cpu().numpy(), rtol=1e-6, atol=1e-6)
        pm = enc.graph_mean(h)
        np.testing.assert_allclose(pm.cpu().numpy(), h.mean(1).cpu().numpy(), rtol=1e-5, atol=1e-6)
