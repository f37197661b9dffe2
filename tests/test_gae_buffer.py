"""4-stream GAE kernel + device rollout buffer against the reference's own GAE method
(tests/golden/gen_gae_golden.py) at FP32 tolerance."""
import importlib
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "gae_golden.npz")


def test_gae4_matches_reference_method():
    buf = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.buffer")
    g = np.load(GOLD)
    dev = "cuda"
    adv = buf.gae4(torch.tensor(g["r"]).to(dev), torch.tensor(g["v"]).to(dev), torch.tensor(g["vn"]).to(dev),
                   torch.tensor(g["done"]).to(dev), 0.99, 0.98)
    np.testing.assert_allclose(adv.cpu().numpy(), g["adv"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(adv.cpu().numpy(), g["adv_local"], rtol=2e-4, atol=2e-5)


def test_gae4_large_block_against_fp64_scan():
    buf = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.buffer")
    gtor = torch.Generator(device="cuda").manual_seed(0)
    T, B = 180, 4099
    r = torch.randn(T, B, 4, device="cuda", generator=gtor)
    v = torch.randn(T, B, 4, device="cuda", generator=gtor)
    vn = torch.randn(T, B, 4, device="cuda", generator=gtor)
    done = (torch.rand(T, B, device="cuda", generator=gtor) < 0.03).float()
    adv = buf.gae4(r, v, vn, done, 0.99, 0.98, normalize=False)
    ref = torch.zeros(T, B, 4, dtype=torch.float64, device="cuda")
    gae = torch.zeros(B, 4, dtype=torch.float64, device="cuda")
    for t in range(T - 1, -1, -1):
        delta = r[t].double() + 0.99 * vn[t].double() - v[t].double()
        gae = delta + 0.99 * 0.98 * gae * (1.0 - done[t].double()).unsqueeze(-1)
        ref[t] = gae
    np.testing.assert_allclose(adv.cpu().numpy(), ref.float().cpu().numpy(), rtol=1e-4, atol=1e-4)
    advn = buf.gae4(r, v, vn, done, 0.99, 0.98)
    refn = (ref - ref.mean(dim=(0, 1))) / (ref.std(dim=(0, 1)) + 1e-5)
    np.testing.assert_allclose(advn.cpu().numpy(), refn.float().cpu().numpy(), rtol=1e-3, atol=1e-4)


def test_rollout_buffer_collects_an_episode():
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    buf = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.buffer")
    B, J, M, E = 64, 6, 6, 2
    d = ins.synthetic_instances(0, B, J, M, E, 4)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(128), 11), J, M)
    mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(128), 12), M)
    ro = rom.Rollout(env, job, mch, greedy=False, seed=1)
    rb = buf.RolloutBuffer(1, env)
    w = ins.random_weights(0, B, 4)
    ro.begin_episode(w)
    wt = torch.as_tensor(w, dtype=torch.float32, device=env.device)
    for s in range(env.N):
        rb.store_pre(env, ro)
        ro.step()
        rb.store_post(env, ro, wt)
    rb.store_bootstrap(torch.zeros(B, 2, device=env.device), torch.zeros(B, 2, device=env.device))
    assert rb.t == env.N and float(rb["done"][-1].sum()) == B and float(rb["done"][:-1].sum()) == 0
    assert torch.equal(rb["job_v_n"][:-1], rb["job_v"][1:])                   # Run.py:448-451
    adv = rb.advantages()
    assert adv.shape == (env.N, B, 4) and torch.isfinite(adv).all()
    np.testing.assert_allclose(adv.mean(dim=(0, 1)).cpu().numpy(), 0.0, atol=1e-4)
    # 122 GB per dense float64 adjacency copy at B = 65,536 in the reference layout; here, everything per env-step:
    assert rb.bytes_per_env_step() < 3200
    dense_ref = 2 * J * M * J * M * 8
    assert rb.bytes_per_env_step() * 6 < dense_ref
