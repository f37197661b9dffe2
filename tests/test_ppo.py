"""Batched PPO update against the reference's own update.

Golden: tests/golden/ppo_golden.npz, produced by tests/golden/gen_ppo_golden.py from the REAL
PPOAlgorithm.global_update_JointActions_GAT_selfCritic on CPU (hidden 32, 2 episodes x 36 steps x 4 envs, K_epochs 2,
minibatches of 36 steps in the recorded SubsetRandomSampler order).  The update here is one batched forward per
minibatch instead of 36 sequential ones; it must land on the same parameters.

Tolerance (floating point, FP32 like the reference): parameters after the 4 Adam steps rtol 2e-3 / atol 5e-4 = half of
one Adam step (a step moves a weight by ~lr = 1e-3 whatever the gradient's size, so rounding in gradients near
lr_eps = 1e-5 is amplified); at most max(2, 1 %) of a tensor's elements may exceed atol 2e-4; losses rtol 1e-3.

The CPU variant swaps the CUDA kernels the update calls (aggregate fwd/bwd, graph mean, grouped BatchNorm fwd/bwd,
GAE) for torch restatements defined in this file, so the host logic is covered without a GPU; the `gpu` variant runs the kernels."""
import importlib
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
ppo = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.ppo")


def dense_to_ell(adj, M):
    """[.., N, N] dense adj[dst, src] (diagonal 1) -> adj_w [.., N, 2] f32, adj_src [.., N] i16 as the env emits them."""
    adj = np.asarray(adj, dtype=np.float64)
    N = adj.shape[-1]
    flat = adj.reshape(-1, N, N)
    w = np.zeros((flat.shape[0], N, 2), dtype=np.float32)
    src = np.full((flat.shape[0], N), -1, dtype=np.int16)
    for v in range(N):
        row = flat[:, v, :].copy()
        row[:, v] = 0
        if v % M != 0:
            w[:, v, 0] = row[:, v - 1]
            row[:, v - 1] = 0
        assert ((row != 0).sum(axis=1) <= 1).all()
        has = (row != 0).any(axis=1)
        col = row.argmax(axis=1)
        w[has, v, 1] = row[has, col[has]]
        src[has, v] = col[has]
    return w.reshape(adj.shape[:-2] + (N, 2)), src.reshape(adj.shape[:-2] + (N,))


def build_batch(dev):
    g = np.load(os.path.join(GOLD, "replay_j6m6_ls_esa.npz"))
    p = np.load(os.path.join(GOLD, "ppo_golden.npz"))
    J, M = int(g["J"]), int(g["M"])
    N, B, EP = J * M, g["t"].shape[0], g["actions"].shape[0]
    T = EP * N
    pre = lambda post, first: np.concatenate([np.concatenate((first[e][None], post[e, :-1]), axis=0) for e in range(EP)], axis=0)
    post = lambda x: x.reshape((T,) + x.shape[2:])
    tfea_post = g["tfea"].reshape(EP, N, B, N, 12)
    tfea0 = g["tfea0"].reshape(EP, B, N, 12)
    cand0 = np.tile(np.arange(J) * M, (EP, B, 1)).astype(np.int16)
    mask0 = np.zeros((EP, B, J), dtype=bool)
    adj_pre, adj_post = pre(g["adj"], g["adj0"]), post(g["adj"])
    aw, asrc = dense_to_ell(adj_pre, M)
    awn, asrcn = dense_to_ell(adj_post, M)
    info = post(g["info"])
    f = lambda x: torch.tensor(np.ascontiguousarray(x), dtype=torch.float32, device=dev)
    bt = dict(
        task_fea=f(pre(tfea_post, tfea0)), adj_w=f(aw), adj_src=torch.tensor(asrc, device=dev),
        candidate=torch.tensor(pre(g["cand"], cand0).astype(np.int32), device=dev),
        job_mask=torch.tensor(pre(g["mask"], mask0).astype(np.uint8), device=dev),
        mach_fea1=f(post(g["mfea1"])), mach_fea2=f(pre(g["mfea2"], g["mfea20"])),
        mach_mask=torch.tensor(p["mach_mask"].astype(np.uint8), device=dev),
        a_job=torch.tensor((post(g["actions"])[..., 0] // M).astype(np.int32), device=dev),
        a_mach=torch.tensor(post(g["actions"])[..., 1].astype(np.int32), device=dev),
        log_a=f(p["log_a"]), m_log_a=f(p["m_log_a"]), job_v=f(p["job_v"]), mch_v=f(p["mch_v"]),
        job_v_n=f(p["job_v_n"]), mch_v_n=f(p["mch_v_n"]),
        r4=f(np.stack((info[..., 2], info[..., 4], info[..., 5], info[..., 3]), axis=-1)),   # mk, pt, tt, it
        done=f(info[..., 1]), rw=f(np.repeat(g["weights"][:, None], N, axis=1).reshape(T, B, 3)),
        task_fea_n=f(post(tfea_post)), adj_w_n=f(awn), adj_src_n=torch.tensor(asrcn, device=dev), mach_fea2_n=f(post(g["mfea2"])),
    )
    return bt, p, (J, M)


# ---- torch restatements of the kernels (CPU variant only) -----------------------------------------------------------
def _t_aggregate(h, adj_w, adj_src, in_scale=None, in_shift=None, relu=False, adj_dst=None):
    B, N, C = h.shape
    prev = torch.cat((torch.zeros_like(h[:, :1]), h[:, :-1]), dim=1)
    has_m = (adj_src >= 0)
    gsrc = torch.gather(h, 1, adj_src.clamp(min=0).long().unsqueeze(-1).expand(-1, -1, C))
    deg = 1.0 + (adj_w[..., 0] != 0).float() + has_m.float()
    return (h + adj_w[..., 0:1] * prev + (adj_w[..., 1:2] * has_m.unsqueeze(-1)) * gsrc) / deg.unsqueeze(-1)


def _t_gae4(r, v, v_next, done, gamma=0.99, lam=0.98, normalize=True):
    T = r.shape[0]
    adv = torch.zeros_like(r)
    gae = torch.zeros_like(r[0])
    delta = r + gamma * v_next - v
    for t in range(T - 1, -1, -1):
        gae = delta[t] + gamma * lam * gae * (1.0 - done[t]).unsqueeze(-1)
        adv[t] = gae
    if normalize:
        flat = adv.reshape(-1, 4)
        adv = (adv - flat.mean(0)) / (flat.std(0) + 1e-5)
    return adv


def _t_bn_forward(x, w, b, eps, groups, relu):
    xs = x.reshape(groups, -1, x.shape[-1])
    var, mean = torch.var_mean(xs, dim=1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    y = (xs - mean) * rstd * w + b
    return (torch.relu(y) if relu else y).reshape(x.shape), mean.squeeze(1), rstd.squeeze(1)


def _t_bn_backward(x, gy, w, b, mean, rstd, groups, relu):
    xs = x.reshape(groups, -1, x.shape[-1])
    R = xs.shape[1]
    xhat = (xs - mean.unsqueeze(1)) * rstd.unsqueeze(1)
    g = gy.reshape(xs.shape)
    if relu:
        g = g * (xhat * w + b > 0)
    sg, sgx = g.sum(dim=1, keepdim=True), (g * xhat).sum(dim=1, keepdim=True)
    dx = (g - (sg + xhat * sgx) / R) * (rstd.unsqueeze(1) * w)
    return dx.reshape(x.shape), sgx.sum(dim=(0, 1)), sg.sum(dim=(0, 1))


def _run_update(dev, monkeypatch=None, max_rows=1 << 21):
    if monkeypatch is not None:
        monkeypatch.setattr(enc, "bn_forward", _t_bn_forward)
        monkeypatch.setattr(enc, "bn_backward", _t_bn_backward)
        monkeypatch.setattr(enc, "aggregate", _t_aggregate)
        monkeypatch.setattr(enc, "graph_mean", lambda h, *a, **k: h.mean(dim=1))
        monkeypatch.setattr(enc, "ell_invert", lambda s: None)
        monkeypatch.setattr(ppo, "gae4", _t_gae4)
    bt, p, (J, M) = build_batch(dev)
    H = int(p["H"])
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 11), J, M, hidden=H, device=dev, trainable=True)
    mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 12), M, hidden=H, device=dev, trainable=True)
    crit = enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(H), 13), J, M, hidden=H, device=dev, trainable=True)
    up = ppo.MAPPOUpdate(job, mch, crit, ppo.PPOConfig(k_epochs=int(p["K_epochs"])), max_rows=max_rows)
    mean, std = up.update(bt, int(p["mini_bs"]), orders=p["orders"])
    np.testing.assert_allclose(mean.cpu().numpy(), p["loss_mean"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(std.cpu().numpy(), p["loss_std"], rtol=2e-2, atol=1e-4)
    worst = 0.0
    for tag, net in (("job", job), ("mch", mch), ("crit", crit)):
        sd = net.state_dict()
        for k in sd:
            if enc._Params.is_parameter(k):
                ref = p["%s/%s" % (tag, k)]
                got = sd[k].cpu().numpy()
                worst = max(worst, float(np.abs(got - ref).max()))
                np.testing.assert_allclose(got, ref, rtol=2e-3, atol=5e-4, err_msg="%s/%s" % (tag, k))
                loose = np.abs(got - ref) > 2e-4 + 2e-3 * np.abs(ref)
                assert loose.sum() <= max(2, 0.01 * loose.size), ("%s/%s" % (tag, k), int(loose.sum()), loose.size)
    return worst


@pytest.mark.parametrize("max_rows", [1 << 21, 7 * 4 * 36])
def test_update_matches_reference_with_torch_kernels(monkeypatch, max_rows):
    """max_rows = 7 steps' worth of node rows walks each 36-step minibatch in 6 gradient-accumulation chunks."""
    _run_update(torch.device("cpu"), monkeypatch, max_rows)


def test_parameters_actually_move():
    """Guards the comparison above against a vacuous pass: the golden parameters differ from the seeded ones."""
    p = np.load(os.path.join(GOLD, "ppo_golden.npz"))
    H = int(p["H"])
    sd = enc.seeded_state_dict(enc.job_actor_keys(H), 11)
    k = "encoder.feature_extract.mlps.0.linears.0.weight"
    assert np.abs(p["job/" + k] - sd[k].numpy()).max() > 1e-3


def test_refresh_twins_rederives_cached_layouts():
    """an inference twin caches re-laid-out copies of some weights; after the optimiser moved the
    shared storage, `refresh_twins()` on the trained network (called by MAPPOUpdate.update) must bring them up to date."""
    H, J, M = 32, 3, 3
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 1), J, M, hidden=H, device="cpu", trainable=True)
    twin = job.inference_twin("fp32")
    W0 = twin.w["o_policy.linears.0.weight"]
    blk = twin._derived("o_policy.W0a", lambda: W0[:, :H])
    assert blk.data_ptr() != W0.data_ptr() and torch.equal(blk, W0[:, :H])
    with torch.no_grad():
        job.w["o_policy.linears.0.weight"].add_(1.0)          # what an optimiser step does: in place, same storage
    assert not torch.equal(blk, W0[:, :H])                     # the cached copy is stale now ...
    job.refresh_twins()
    assert torch.equal(blk, W0[:, :H])                         # ... and current again, in the same tensor (graph-safe)
    del twin, blk, W0
    import gc

    gc.collect()
    job.refresh_twins()                                        # dead twins are dropped, not dereferenced
    assert job._twins == []


@pytest.mark.gpu
@pytest.mark.parametrize("max_rows", [1 << 21, 7 * 4 * 36])
def test_update_matches_reference_on_device(max_rows):
    _run_update(torch.device("cuda", 0), None, max_rows)


@pytest.mark.gpu
def test_aggregate_backward_matches_autograd_of_dense_product():
    dev = torch.device("cuda", 0)
    bt, p, (J, M) = build_batch(dev)
    N = J * M
    aw, asrc = bt["adj_w_n"][40:60].reshape(-1, N, 2).contiguous(), bt["adj_src_n"][40:60].reshape(-1, N).contiguous()
    for C in (12, 32, 128):
        h = torch.randn(aw.shape[0], N, C, device=dev, requires_grad=True)
        g = torch.randn_like(h)
        out = enc.aggregate(h, aw, asrc)
        out.backward(g)
        h2 = h.detach().clone().requires_grad_(True)
        ref = _t_aggregate(h2.double(), aw.double(), asrc)
        ref.backward(g.double())
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().float().cpu().numpy(), rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(h.grad.cpu().numpy(), h2.grad.cpu().numpy(), rtol=1e-5, atol=1e-4)
        pm = enc.graph_mean(h)
        pm.backward(torch.ones_like(pm))


@pytest.mark.gpu
def test_collect_feeds_update_consistently():
    """collect() -> update() plumbing on a live rollout: with every minibatch equal to one episode in temporal order
    the batched re-forward sees exactly what the rollout saw (item 0 gets `_input`, item i the machine embedding of
    step i-1), so before the first optimiser step both importance ratios are 1; afterwards parameters have moved and
    losses are finite.  Also the next-state bookkeeping of Run.py:448-475."""
    dev = torch.device("cuda", 0)
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    B, J, M, E, H = 64, 6, 6, 2, 32
    N = J * M
    d = ins.synthetic_instances(0, B, J, M, E, 5)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 1), J, M, hidden=H, trainable=True)
    mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 2), M, hidden=H, trainable=True)
    crit = enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(H), 3), J, M, hidden=H, trainable=True)
    ro = rom.Rollout(env, job, mch, greedy=False, seed=9)
    ws = [ins.random_weights(0, B, 100 + e) for e in range(2)]
    bt = ppo.collect(ro, ws)
    T = 2 * N
    buf = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.buffer")
    assert isinstance(bt, buf.RolloutBuffer) and bt.t == T                       # the tested buffer IS the training path's
    every = torch.arange(T, device=dev)
    assert bt.obs("task_fea", every).shape == (T, B, N, 12)
    assert bt.slots["task_fea"].shape[0] == 2 * (N + 1)                         # each observation stored once
    assert bt.bytes_per_env_step() < 2800
    done = bt["done"].reshape(2, N, B)
    assert bool((done[:, :-1] == 0).all()) and bool((done[:, -1] == 1).all())
    for e in range(2):                                         # in-episode next values are the next step's values
        sl = slice(e * N, e * N + N - 1)
        assert torch.equal(bt["job_v_n"][sl], bt["job_v"][e * N + 1:e * N + N])
        assert torch.equal(bt["mch_v_n"][sl], bt["mch_v"][e * N + 1:e * N + N])
        assert torch.equal(bt.obs("task_fea", every[sl], nxt=True), bt.obs("task_fea", every[e * N + 1:e * N + N]))
    # the terminal observation of episode 0 is its own slot, not the first observation of episode 1
    assert not torch.equal(bt.obs("task_fea", every[N - 1:N], nxt=True), bt.obs("task_fea", every[N:N + 1]))
    assert bool((bt["a_job"] >= 0).all()) and bool((bt["a_job"] < J).all())

    up = ppo.MAPPOUpdate(job, mch, crit, ppo.PPOConfig(k_epochs=1))
    seen = []
    orig = torch.exp

    def spy(x):
        seen.append(x.detach().clone())
        return orig(x)

    before = {k: v.clone() for k, v in job.state_dict().items()}
    ppo.torch.exp = spy
    try:
        mean, _ = up.update(bt, N, orders=[list(range(T))])
    finally:
        ppo.torch.exp = orig
    # first minibatch = episode 0 in order: log-ratio of both actors is ~0 everywhere (FP32 re-evaluation noise)
    assert float(seen[0].abs().max()) < 2e-4 and float(seen[1].abs().max()) < 2e-4, (float(seen[0].abs().max()), float(seen[1].abs().max()))
    assert bool(torch.isfinite(mean).all())
    after = job.state_dict()
    moved = max(float((after[k] - before[k]).abs().max()) for k in before if enc._Params.is_parameter(k))
    assert moved > 1e-4


@pytest.mark.gpu
def test_train_iteration_with_tf32_twins_refreshes_them_and_starts_at_ratio_one():
    """(high) `train_iteration` leaves the TF32 rollout twins consistent with the updated weights;
    (medium) the behaviour log-probabilities of a TF32-twin rollout are re-evaluated on the update's own path, so the
    importance ratios of the first minibatch start at 1 (|log ratio| < 2e-4) instead of a few percent off."""
    dev = torch.device("cuda", 0)
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    B, J, M, E, H = 256, 6, 6, 2, 128
    N = J * M
    d = ins.synthetic_instances(0, B, J, M, E, 5)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 1), J, M, hidden=H, trainable=True)
    mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 2), M, hidden=H, trainable=True)
    crit = enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(H), 3), J, M, hidden=H, trainable=True)
    tj, tm = job.inference_twin("tf32"), mch.inference_twin("tf32")
    ro = rom.Rollout(env, tj, tm, greedy=False, seed=9)
    ws = [ins.random_weights(0, B, 100)]
    bt = ppo.collect(ro, ws)
    raw_la = bt["log_a"].clone()
    up = ppo.MAPPOUpdate(job, mch, crit, ppo.PPOConfig(k_epochs=1))
    up.recompute_old_logp(bt)
    drift = float((bt["log_a"] - raw_la).abs().max())
    assert 0 < drift < 0.2, drift                              # TF32 vs FP32 policies differ, by a bounded amount
    seen = []
    orig = torch.exp

    def spy(x):
        seen.append(x.detach().clone())
        return orig(x)

    ppo.torch.exp = spy
    try:
        up.update(bt, N, orders=[list(range(N))])
    finally:
        ppo.torch.exp = orig
    assert float(seen[0].abs().max()) < 2e-4 and float(seen[1].abs().max()) < 2e-4
    for twin in (tj, tm):                                      # every cached layout equals its source after the update
        assert twin._dcache
        for t, fn in twin._dcache.values():
            assert torch.equal(t, fn().contiguous())
    # and a second whole iteration runs on the refreshed twins
    mean, _ = ppo.train_iteration(ro, up, ws)
    assert bool(torch.isfinite(mean).all())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 700, 128), (5, 1031, 128), (3, 64, 32), (36, 24, 32)])
@pytest.mark.parametrize("relu", [False, True])
def test_grouped_batchnorm_kernels_match_torch_autograd(shape, relu):
    dev = torch.device("cuda", 0)
    G, R, C = shape
    torch.manual_seed(G * 1000 + R)
    x = (torch.randn(G * R, C, device=dev) * 2 + 0.5).requires_grad_(True)
    w = (torch.rand(C, device=dev) + 0.5).requires_grad_(True)
    b = (torch.randn(C, device=dev) * 0.3).requires_grad_(True)
    gy = torch.randn(G * R, C, device=dev)
    y = enc._bn_train(x, w, b, groups=G, relu=relu)
    y.backward(gy)
    x2, w2, b2 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    xs = x2.reshape(G, R, C)
    var, mean = torch.var_mean(xs, dim=1, unbiased=False, keepdim=True)
    ref = ((xs - mean) * torch.rsqrt(var + 1e-5) * w2 + b2).reshape(G * R, C)
    if relu:
        ref = torch.relu(ref)
    ref.backward(gy.double())
    close = lambda a, r, name: np.testing.assert_allclose(a.detach().cpu().numpy(), r.detach().float().cpu().numpy(), rtol=2e-4,
                                                          atol=2e-4, err_msg=name)
    close(y, ref, "y"); close(x.grad, x2.grad, "dx"); close(w.grad, w2.grad, "dgamma"); close(b.grad, b2.grad, "dbeta")
    with torch.no_grad():
        close(enc._bn_train(x, w, b, groups=G, relu=relu), ref, "y (no grad)")


@pytest.mark.gpu
def test_update_with_tcgen05_encoder_gemms_tracks_the_fp32_update():
    """PPOConfig.encoder_tf32: the graph encoders' Linear layers (forward, input gradient, weight gradient) on the
    hand-written tcgen05 kernels.  Same buffer, same initial weights, one epoch of two minibatches: the losses of the
    first minibatch agree to TF32 tolerance, and the accumulated gradients of the encoder weights point the same way
    (cosine > 0.99 for the actor, > 0.95 for the critic -- 10-bit-mantissa operands through 6 GEMMs and 8 batch norms, then
    backwards)."""
    dev = torch.device("cuda", 0)
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
    B, J, M, E, H = 256, 6, 6, 2, 128
    N = J * M
    d = ins.synthetic_instances(0, B, J, M, E, 5)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()

    def nets():
        return (enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 1), J, M, hidden=H, trainable=True),
                enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 2), M, hidden=H, trainable=True),
                enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(H), 3), J, M, hidden=H, trainable=True))

    job, mch, crit = nets()
    ro = rom.Rollout(env, job, mch, greedy=False, seed=9)
    bt = ppo.collect(ro, [ins.random_weights(0, B, 100)])
    res = {}
    for tf in (False, True):
        j, m, c = nets()
        up = ppo.MAPPOUpdate(j, m, c, ppo.PPOConfig(k_epochs=1, encoder_tf32=tf))
        grads = {}
        step_j, step_c = up.opt_job.step, up.opt_critic.step

        def spy_j(*a, _g=grads, _j=j, **k):
            _g.setdefault("job", [p.grad.detach().clone() for p in _j.parameters() if p.grad is not None])
            return step_j(*a, **k)

        def spy_c(*a, _g=grads, _c=c, **k):
            _g.setdefault("critic", [p.grad.detach().clone() for p in _c.parameters() if p.grad is not None])
            return step_c(*a, **k)

        up.opt_job.step, up.opt_critic.step = spy_j, spy_c
        mean, _ = up.update(bt, N // 2, orders=[list(range(N))])
        res[tf] = (mean, grads, [p.detach().clone() for p in j.parameters()])
    m0, g0, p0 = res[False]
    m1, g1, p1 = res[True]
    assert bool(torch.isfinite(m1).all())
    np.testing.assert_allclose(m1.cpu().numpy(), m0.cpu().numpy(), rtol=5e-2, atol=5e-3)
    for net in ("job", "critic"):
        a = torch.cat([t.reshape(-1) for t in g0[net]]).double()
        b = torch.cat([t.reshape(-1) for t in g1[net]]).double()
        cos = float((a @ b) / (a.norm() * b.norm()))
        # the actor gradient stays above 0.99 on every buffer tried; the critic's (value regression through 6 GEMMs and 8
        # batch norms) depends on the sampled trajectories: 0.995 on the buffer torch.multinomial drew, 0.96 on the one the
        # selection kernel draws from the same seed
        assert cos > (0.99 if net == "job" else 0.95), (net, cos)
    moved = max(float((x - y).abs().max()) for x, y in zip(p1, [p.detach() for p in nets()[0].parameters()]))
    assert moved > 1e-4
