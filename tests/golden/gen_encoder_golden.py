"""Golden vectors for the encoder side, produced by the REAL reference modules on CPU:

    model/actor_critic.py  Operation_Actor_JointAction_selfCritic, Machine_Actor_JointAction_selfGAT_selfCritic

with weights from `encoder.seeded_state_dict` (the same deterministic call re-creates them on the GPU box, so no
checkpoint is committed) and inputs taken from the reference env dump tests/golden/replay_j6m6_ls_esa.npz.
Also writes the key -> shape table of the shipped checkpoints (tester/IoTJ_MAPPO/*.pth) for the load-compat test.

Run in the build container only:  python tests/golden/gen_encoder_golden.py
"""
import contextlib
import importlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")


def load_models():
    rh.load_reference()
    td = types.ModuleType("trainer.train_device")
    td.device = torch.device("cpu")
    sys.modules["trainer.train_device"] = td
    fk = types.ModuleType("trainer.fig_kpi")
    fk.get_GPU_usage = lambda *a, **k: None
    fk.result_box_plot = lambda *a, **k: None
    sys.modules["trainer.fig_kpi"] = fk
    with contextlib.redirect_stdout(io.StringIO()):
        from model import actor_critic
        from model.gcn_mlp import g_pool_cal
    return actor_critic, g_pool_cal


def configs(J, M, B, H):
    return {"n_job": J, "n_machine": M, "env_batch": B, "GAMMA": 0.99, "LAMDA": 0.98, "epsilon": 0.2, "ENTROPY_BETA": 0.01,
            "gcn_layer": 3, "mlp_fea_extract_layer": 3, "gcn_input_dim": 12, "gcn_hidden_dim": H, "learn_eps": False,
            "neighbor_pooling_type": "average", "mlp_actor_layer": 3, "mlp_critic_layer": 3, "critic_input_dim": H,
            "critic_hidden_dim": H, "use_orthogonal": False, "machine_hidden_dim": H}


def main():
    ac, g_pool_cal = load_models()
    g = np.load(os.path.join(HERE, "replay_j6m6_ls_esa.npz"))
    J, M = int(g["J"]), int(g["M"])
    N, B = J * M, g["t"].shape[0]
    out = {}
    for H in (128, 32):
        cfg = configs(J, M, B, H)
        with contextlib.redirect_stdout(io.StringIO()):
            job = ac.Operation_Actor_JointAction_selfCritic(cfg)
            mch = ac.Machine_Actor_JointAction_selfGAT_selfCritic(cfg)
        job.load_state_dict(enc.seeded_state_dict(enc.job_actor_keys(H), 11), strict=True)
        mch.load_state_dict(enc.seeded_state_dict(enc.machine_actor_keys(H), 12), strict=True)
        gp = g_pool_cal("average", B, N, torch.device("cpu"))
        h_m = None
        for tag, s in (("init", -1), ("s05", 5), ("s20", 20), ("s34", 34)):
            if s < 0:
                tfea, adj = g["tfea0"][0], g["adj0"][0].astype(np.float64)
                cand = np.tile(np.arange(J) * M, (B, 1)); mask = np.zeros((B, J), dtype=bool)
                mfea2 = g["mfea20"][0]
            else:
                tfea, adj = g["tfea"][0, s], g["adj"][0, s].astype(np.float64)
                cand, mask, mfea2 = g["cand"][0, s].astype(np.int64), g["mask"][0, s], g["mfea2"][0, s]
            with torch.no_grad():
                ti, ai, la, prob, hgo, jv = job(tfea, gp, None, adj, cand, h_m, torch.tensor(mask), use_greedy=True)
                # candidate-machine features of the greedily chosen op come from the env dump's next step when the
                # action matches; for the fixture any valid op works: use the recorded action of step s+1
                sn = s + 1
                mfea1 = g["mfea1"][0, sn]
                op = g["actions"][0, sn][:, 0]
                mmask = torch.tensor(g["t"][np.arange(B), op] < 0)[:, None, :]
                mp, hp, mv = mch(mfea1, mfea2, hgo, mmask)
            k = "H%d_%s_" % (H, tag)
            out[k + "prob"] = prob.numpy(); out[k + "pooled"] = hgo.numpy(); out[k + "job_v"] = jv.numpy()
            out[k + "task_index"] = ti.numpy(); out[k + "hgm_in"] = np.zeros((0,)) if h_m is None else h_m.numpy()
            out[k + "mch_prob"] = mp.numpy(); out[k + "mch_pooled"] = hp.numpy(); out[k + "mch_v"] = mv.numpy()
            out[k + "mmask"] = mmask.numpy()[:, 0, :]
            out[k + "step"] = np.array(s)
            h_m = hp
    np.savez_compressed(os.path.join(HERE, "encoder_golden.npz"), **out)
    # key tables of the shipped checkpoints
    tab = {}
    for f in ("PPO_operation_actor_J6M6E2_1000.pth", "PPO_machine_actor_J6M6E2_1000.pth"):
        sd = torch.load(os.path.join(rh.REF_ROOT, "tester/IoTJ_MAPPO", f), map_location="cpu")
        tab[f] = {k: list(v.shape) for k, v in sd.items()}
    json.dump(tab, open(os.path.join(HERE, "shipped_checkpoint_keys.json"), "w"), indent=0)
    print("wrote encoder_golden.npz (%d KB)" % (os.path.getsize(os.path.join(HERE, "encoder_golden.npz")) // 1024))


if __name__ == "__main__":
    main()
