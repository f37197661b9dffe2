"""On-device rollout controller: greedy rollouts with the device actors are replayed on the CPU
oracle from the recorded actions; the CUDA-graph replay path must produce the same episode as the eager path."""
import importlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _setup(B, precision="fp32"):
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    J, M, E = 6, 6, 2
    d = ins.synthetic_instances(0, B, J, M, E, 21)
    w = ins.random_weights(0, B, 21)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(128), 11), J, M, precision=precision)
    mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(128), 12), M)
    return env, job, mch, d, w


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_greedy_rollout_replays_on_the_oracle(precision):
    from oracle.mtfjsp_oracle import OracleEnv

    ro_mod = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    B = 128
    env, job, mch, d, w = _setup(B, precision)
    ro = ro_mod.Rollout(env, job, mch, greedy=True)
    ro.begin_episode(w)
    ora = OracleEnv(B, 6, 6, 2)
    ora.load(d["t"], d["p"], d["transT"], d["edge"])
    ora.scaler_init()
    ora.reset(w)
    for s in range(env.N):
        ro.step()
        op, mc = env.op.cpu().numpy(), env.mach.cpu().numpy()
        assert not env.invalid.cpu().numpy().any()          # masked policies only emit valid actions
        r5, s4, done, inv = ora.step(op, mc)
        assert not inv.any()
        np.testing.assert_array_equal(env.reward5.cpu().numpy(), r5)
        np.testing.assert_array_equal(env.scaled4.cpu().numpy(), s4)
        np.testing.assert_array_equal(env.task_fea.cpu().numpy(), ora.obs(1)["task_fea"].astype(np.float32))
    assert env.done.cpu().numpy().all()
    np.testing.assert_array_equal(env.costs().cpu().numpy(), ora.costs())


def test_cuda_graph_replay_equals_eager():
    ro_mod = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    B = 96
    env, job, mch, d, w = _setup(B)
    eager = ro_mod.Rollout(env, job, mch, greedy=True).run_episode(w).clone()
    acts_e = env.op.clone()
    graph_ro = ro_mod.Rollout(env, job, mch, greedy=True, use_cuda_graph=True)
    c1 = graph_ro.run_episode(w).clone()
    c2 = graph_ro.run_episode(w).clone()                    # second episode is pure graph replays
    assert env.done.cpu().numpy().all()
    np.testing.assert_array_equal(c1.cpu().numpy(), eager.cpu().numpy())
    np.testing.assert_array_equal(c2.cpu().numpy(), eager.cpu().numpy())
    np.testing.assert_array_equal(env.op.cpu().numpy(), acts_e.cpu().numpy())


def test_sampled_rollout_finishes_with_valid_schedules():
    ro_mod = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    B = 256
    env, job, mch, d, w = _setup(B)
    ro = ro_mod.Rollout(env, job, mch, greedy=False, seed=3)
    ro.begin_episode(w)
    ninv = 0
    for s in range(env.N):
        ro.step()
        ninv += int(env.invalid.sum().item())
    assert ninv == 0 and int(env.done.sum().item()) == B
    st = {k: v.cpu().numpy() for k, v in env.export_state().items()}
    dur = np.take_along_axis(d["t"], st["mach"][:, :, None].astype(np.int64), axis=2)[:, :, 0]
    assert (dur > 0).all()
    np.testing.assert_array_equal(st["ft"], st["st"] + dur)


@pytest.mark.parametrize("size", [(10, 10, 3, 96), (20, 6, 3, 64), (30, 20, 5, 8)])
def test_tf32_actors_at_other_sizes(size):
    """Both actors on the tcgen05 path at sizes with other tile shapes: N = 100 (aggregation in the layer's epilogue with
    100-row tiles), N = 120 and N = 600 (separate aggregation kernel: more than 128 nodes), heads with 10 / 20 / 30 rows
    per env, trunk with 6 / 10 / 20 machines.  First-step distributions against the FP32 actors, then a whole sampled
    episode under CUDA-graph replay: only valid actions, every env finishes."""
    J, M, E, B = size
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    ro_mod = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
    d = ins.synthetic_instances(0, B, J, M, E, 77)
    w = ins.random_weights(0, B, 77)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    env.reset(w)
    env.obs(1)
    sdj, sdm = enc.seeded_state_dict(enc.job_actor_keys(128), 21), enc.seeded_state_dict(enc.machine_actor_keys(128), 22)
    with torch.no_grad():
        probs = {}
        for prec in ("fp32", "tf32"):
            job = enc.JobActor(sdj, J, M, precision=prec)
            mch = enc.MachineActor(sdm, M, precision=prec)
            pj, pooled, _ = job.evaluate(env.task_fea, env.adj_w, env.adj_src, env.candidate, None, env.job_mask)
            op = env.candidate.long().gather(1, pj.argmax(dim=-1, keepdim=True)).squeeze(-1).to(torch.int32)
            m1, mmask = env.mfea1(op if prec == "fp32" else probs["op"])
            pm, _, _ = mch.forward(m1, env.mach_fea, pooled, mmask)
            probs[prec] = (pj, pm)
            probs.setdefault("op", op)
    np.testing.assert_allclose(probs["tf32"][0].cpu().numpy(), probs["fp32"][0].cpu().numpy(), rtol=5e-2, atol=5e-3)
    np.testing.assert_allclose(probs["tf32"][1].cpu().numpy(), probs["fp32"][1].cpu().numpy(), rtol=5e-2, atol=5e-3)
    ro = ro_mod.Rollout(env, enc.JobActor(sdj, J, M, precision="tf32"), enc.MachineActor(sdm, M, precision="tf32"), greedy=False,
                        use_cuda_graph=True, seed=5)
    ro.begin_episode(w)
    ninv = 0
    for s in range(env.N):
        ro.step()
        ninv += int(env.invalid.sum().item())
    assert ninv == 0 and int(env.done.sum().item()) == B
