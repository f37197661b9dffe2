"""Pins the CPU restatement (oracle/mtfjsp_oracle.c) to the reference.

(1) the reference's shipped result rows: results/test_results/Real_{MK,PT,TT,IT}_J6_M6_E2_Seed3_Weight442.csv,
    10 deterministic rule rows x 100 test instances x 4 costs = 4000 known answers, bit-exact;
(2) per-step dumps of the reference Parallel_env + job-mask rule (tests/golden/gen_golden.py), bit-exact
    on every output: adjacency, 12-wide task features, 8- and 6-wide machine features, rewards, scaled
    rewards, done, job mask, candidates, machine assignment, start / finish times, machine routes.
"""
import glob
import os

import numpy as np
import pytest

from oracle.mtfjsp_oracle import OracleEnv, np_sum


def test_np_sum_restatement_matches_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(1, 160)) + [255, 256, 257, 600, 601, 1000, 4097]:
        for _ in range(5):
            a = rng.random(n) * 1000.0
            assert np_sum(a) == float(np.sum(a)), n


def test_pdr_rows_match_shipped_csv(golden_dir):
    g = np.load(os.path.join(golden_dir, "pdr_golden.npz"))
    gold, ops, mch = g["gold"], g["ops"], g["mch"]
    R, S, N = ops.shape
    env = OracleEnv(S, 6, 6, 2, left_shift=False)  # tester/pdrs.py:669 runs the rules without left shift
    env.load(g["t"], g["p"], g["transT"], g["edge"])
    env.scaler_init()
    w = np.tile(np.array([0.4, 0.4, 0.2]), (S, 1))  # reset(Random_weight_type="eval")
    for r in range(R):
        env.reset(w)
        for s in range(N):
            r5, s4, done, inv = env.step(ops[r, :, s], mch[r, :, s])
            assert not inv.any()
        assert done.all()
        np.testing.assert_array_equal(env.costs(), gold[r], err_msg=str(g["rule_names"][r]))


def _replay_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "replay_*.npz")))


def check_replay(make_env, path, exact=True):
    """Shared by the oracle test (here) and the CUDA parity test (tests/test_cuda_parity.py)."""
    g = np.load(path)
    J, M, E = int(g["J"]), int(g["M"]), int(g["E"])
    N = J * M
    t = g["t"]
    B = t.shape[0]
    env = make_env(B, J, M, E, bool(g["left_shift"]))
    env.load(t, g["p"], g["transT"], g["edge"])
    env.scaler_init()
    mm = int(g["mask_mode"])
    eq = np.testing.assert_array_equal
    # large fixtures keep the bulky dumps (observation arrays, schedule state) at g["obs_steps"] only
    kept = {int(s): k for k, s in enumerate(g["obs_steps"])} if "obs_steps" in g.files else None
    for ep in range(g["weights"].shape[0]):
        env.reset(g["weights"][ep])
        env.scaler_reset()
        ob = env.obs(mm)
        eq(ob["task_fea"].reshape(-1, 12), g["tfea0"][ep])
        eq(ob["mach_fea"], g["mfea20"][ep])
        eq(env.dense_adj(), g["adj0"][ep].astype(np.float64))
        for s in range(N):
            act = g["actions"][ep, s].astype(np.int32)
            m1, mmask = env.mfea1(act[:, 0])
            eq(m1, g["mfea1"][ep, s])
            eq(mmask.astype(bool), t[np.arange(B), act[:, 0]] < 0)
            r5, s4, done, inv = env.step(act[:, 0], act[:, 1])
            assert not inv.any()
            info = g["info"][ep, s]
            eq(r5[:, 0], info[:, 0])
            eq(done.astype(np.float64), info[:, 1])
            eq(s4, info[:, 2:6])
            ob = env.obs(mm)
            if mm == 1:
                eq(ob["job_mask"].astype(bool), g["mask"][ep, s])
            eq(ob["candidate"], g["cand"][ep, s])
            if kept is not None and s not in kept:
                continue
            k = s if kept is None else kept[s]
            eq(ob["task_fea"].reshape(-1, 12), g["tfea"][ep, k])
            eq(ob["mach_fea"], g["mfea2"][ep, k])
            eq(env.dense_adj(), g["adj"][ep, k].astype(np.float64))
            st = env.export_state()
            eq(st["mach"], g["mach"][ep, k])
            eq(st["st"], g["st"][ep, k])
            eq(st["ft"], g["ft"][ep, k])
            eq(st["routes"], g["routes"][ep, k])
        eq(env.costs(), g["costs"][ep])


@pytest.mark.parametrize("name", [os.path.basename(p) for p in _replay_files(os.path.join(os.path.dirname(__file__), "golden"))])
def test_replay_matches_reference_dump(golden_dir, name):
    check_replay(lambda B, J, M, E, ls: OracleEnv(B, J, M, E, left_shift=ls), os.path.join(golden_dir, name))


def test_invalid_actions_are_flagged_and_leave_state_untouched():
    from importlib import import_module

    ins = import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    d = ins.synthetic_instances(0, 4, 3, 3, 1, 5)
    env = OracleEnv(4, 3, 3, 1)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    env.reset(np.full((4, 3), 1 / 3))
    before = env.export_state()
    feas = np.argmax(d["t"][:, 0] >= 0, axis=1)
    infeas = np.argmax(d["t"][:, 1] < 0, axis=1)
    # op 1 before op 0 (precedence), out-of-range op, out-of-range machine
    for op, mc in ((np.full(4, 1), feas), (np.full(4, 99), feas), (np.zeros(4), np.full(4, 7))):
        _, _, _, inv = env.step(op, mc)
        assert inv.all()
    after = env.export_state()
    for k in before:
        np.testing.assert_array_equal(before[k], after[k])
    _, _, _, inv = env.step(np.zeros(4), feas)
    assert not inv.any()
    _, _, _, inv = env.step(np.zeros(4), feas)  # already scheduled
    assert inv.all()
    has_neg = (d["t"][:, 1] < 0).any(axis=1)
    _, _, _, inv = env.step(np.ones(4), infeas)  # infeasible machine where one exists
    np.testing.assert_array_equal(inv.astype(bool), has_neg)
