"""The drop-in `Parallel_env` (reference surface: trainer/parallel_env.py:19-282) driven exactly as Run.py drives
the reference one, checked against the reference's own per-step dumps (tests/golden/replay_*.npz)."""
import glob
import importlib
import os
import random

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
eq = np.testing.assert_array_equal


def _args(J, M, E, B):
    return {"n_job": J, "n_machine": M, "n_edge": E, "env_batch": B, "m_scaling": 1,
            "reward_scaling": {"scaling_divisor": 1}, "GAMMA": 0.99, "gcn_input_dim": 12,
            "weight_mk": 0.4, "weight_ec": 0.4, "weight_tt": 0.2}


def test_reward_weight_draws_follow_the_reference_random_stream():
    pe = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.parallel_env")
    random.seed(123)
    mine = [pe.draw_reward_weights("01", {}) for _ in range(5)]
    random.seed(123)
    for w in mine:  # singlestep.py:1255-1259
        lst = [random.uniform(0, 1) for _ in range(3)]
        ref = np.array(lst)
        ref = ref / np.sum(ref, axis=-1)
        eq(w, ref)
    random.seed(7)
    mine = pe.draw_reward_weights("0.1", {})
    random.seed(7)
    nums = [round(random.uniform(0, 1), 1) for _ in range(3)]
    eq(mine, np.array([round(n / sum(nums), 1) for n in nums]))
    eq(pe.draw_reward_weights("eval", {"weight_mk": 0.4, "weight_ec": 0.4, "weight_tt": 0.2}), np.array([0.4, 0.4, 0.2]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [os.path.basename(p) for p in sorted(glob.glob(os.path.join(GOLD, "replay_*_ls_*.npz")))])
def test_parallel_env_reproduces_reference_dump(name):
    torch = pytest.importorskip("torch")
    pe_mod = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.parallel_env")
    g = np.load(os.path.join(GOLD, name))
    J, M, E = int(g["J"]), int(g["M"]), int(g["E"])
    N, B = J * M, g["t"].shape[0]
    mm = int(g["mask_mode"])
    paral_env = pe_mod.Parallel_env(_args(J, M, E, B))
    paral_env.get_batch({"t": torch.tensor(g["t"]), "p": torch.tensor(g["p"]), "transT": torch.tensor(g["transT"]),
                         "edge": torch.tensor(g["edge"])})
    paral_env.init_RewardScaling_sameBATCH(shape=4)
    jm = pe_mod.JobMask(paral_env, use_esa=(mm == 1))
    # large fixtures keep the observation / schedule dumps at g["obs_steps"] only (tests/golden/gen_golden.py)
    kept = {int(s): k for k, s in enumerate(g["obs_steps"])} if "obs_steps" in g.files else None
    for ep in range(g["weights"].shape[0]):
        adj, mfea2, tfea = paral_env.init_DGFJSPEnv_state0(weights=g["weights"][ep])
        eq(adj, g["adj0"][ep].astype(np.float64)); eq(mfea2, g["mfea20"][ep]); eq(tfea, g["tfea0"][ep])
        assert adj.dtype == np.float64 and tfea.shape == (B * N, 12) and mfea2.shape == (B, M, 8)
        for rs in paral_env.paral_Rscaling_instance:  # Run.py:283-284
            rs.reset()
        assert paral_env.paral_env_DG[0].G.nodes[1]["finish_time"] is None
        for s in range(N):
            act = g["actions"][ep, s]
            task_index = torch.tensor(act[:, 0].astype(np.int64))
            m_mask = torch.tensor(g["t"][np.arange(B), act[:, 0]] < 0)[:, None, :]
            mfea1 = paral_env.cal_cur_task_machine_feature(task_index, m_mask, tfea)
            eq(mfea1, g["mfea1"][ep, s])
            joint_actions = list(zip(act[:, 0].tolist(), act[:, 1].tolist()))  # Run.py:411
            adj, oenv_info, mfea2, tfea = paral_env.DGFJSPEnv_paral_step(joint_actions)
            eq(np.array(oenv_info, dtype=np.float64), g["info"][ep, s])
            cand, mask = jm.esa_update_chosenTaskID_CandidateTaskIDx_JobMask(paral_env, None, 1)
            eq(cand, g["cand"][ep, s])
            if mm == 1:
                eq(mask.cpu().numpy(), g["mask"][ep, s])
            if kept is not None and s not in kept:
                continue
            k = s if kept is None else kept[s]
            eq(adj, g["adj"][ep, k].astype(np.float64)); eq(mfea2, g["mfea2"][ep, k]); eq(tfea, g["tfea"][ep, k])
            # what algorithm/ppo_algorithm.py:271-273 reads out of the env graphs
            b0 = 0
            ft0 = [paral_env.paral_env_DG[b0].G.nodes[i + 1]["finish_time"] for i in range(N)]
            ref_ft = [g["ft"][ep, k, b0, i] if g["mach"][ep, k, b0, i] >= 0 else None for i in range(N)]
            assert ft0 == ref_ft
        envs = paral_env.paral_env_DG  # Run.py:632-633
        costs = np.array([[e.makespan_previous_step, e.total_e1_previous_step / N, e.trans_t_previous_step,
                           e.idle_t_previous_step] for e in envs])
        eq(costs, g["costs"][ep])
        eq(np.array([e.reward_random_weight for e in envs]), g["weights"][ep])
        for e in envs:  # Run.py:660
            e.reset()
        paral_env.reset_data()
        assert paral_env.paral_env_DG == []


@pytest.mark.gpu
def test_ell_compat_mode_carries_the_same_observation_without_leaving_the_device():
    """compat="ell": adjacency as (adj_w, adj_src) device tensors; rebuilt block-sparse matrix == the dense mode's."""
    torch = pytest.importorskip("torch")
    pe_mod = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.parallel_env")
    compat = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.compat")
    g = np.load(os.path.join(GOLD, "replay_j6m6_ls_esa.npz"))
    J, M, E = int(g["J"]), int(g["M"]), int(g["E"])
    N, B = J * M, g["t"].shape[0]
    envs = {}
    for mode in ("dense", "ell"):
        pe = pe_mod.Parallel_env(_args(J, M, E, B), compat=mode)
        pe.get_batch({"t": torch.tensor(g["t"]), "p": torch.tensor(g["p"]), "transT": torch.tensor(g["transT"]),
                      "edge": torch.tensor(g["edge"])})
        pe.init_RewardScaling_sameBATCH(shape=4)
        envs[mode] = (pe, pe.init_DGFJSPEnv_state0(weights=g["weights"][0]))
    for s in range(N):
        act = g["actions"][0, s]
        ja = list(zip(act[:, 0].tolist(), act[:, 1].tolist()))
        adj_d, info_d, mf_d, tf_d = envs["dense"][0].DGFJSPEnv_paral_step(ja)
        (aw, asrc), info_e, mf_e, tf_e = envs["ell"][0].DGFJSPEnv_paral_step(ja)
        assert aw.is_cuda and asrc.is_cuda and mf_e.is_cuda and tf_e.is_cuda
        eq(np.asarray(info_d, dtype=np.float64), info_e)
        eq(mf_d, mf_e.cpu().numpy()); eq(tf_d, tf_e.cpu().numpy())
        blk = compat.ell_to_block_sparse(aw, asrc).to_dense().cpu().numpy()
        want = np.zeros((B * N, B * N))
        for b in range(B):
            want[b * N:(b + 1) * N, b * N:(b + 1) * N] = adj_d[b]
        eq(blk, want)
