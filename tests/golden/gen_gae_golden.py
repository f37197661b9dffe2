"""Golden vectors for the 4-stream GAE from the REAL reference method
algorithm/ppo_algorithm.py:488-536 `PPOAlgorithm.separate_cal_4_reward_GAE` (called unbound on a stub carrying GAMMA /
LAMDA, so no networks are built).  Run in the build container only:  python tests/golden/gen_gae_golden.py"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

rh.load_reference()
td = types.ModuleType("trainer.train_device"); td.device = torch.device("cpu"); sys.modules["trainer.train_device"] = td
fk = types.ModuleType("trainer.fig_kpi"); fk.get_GPU_usage = lambda *a, **k: None; fk.result_box_plot = lambda *a, **k: None
sys.modules["trainer.fig_kpi"] = fk
with contextlib.redirect_stdout(io.StringIO()):
    from algorithm.ppo_algorithm import PPOAlgorithm

rs = np.random.RandomState(5)
T, B = 36, 7
r = rs.standard_normal((T, B, 4)).astype(np.float32)
v = rs.standard_normal((T, B, 4)).astype(np.float32)
vn = rs.standard_normal((T, B, 4)).astype(np.float32)
done = np.zeros((T, B), dtype=np.float32)
done[17] = 1.0  # two 18-step episodes back to back in the buffer
done[35] = 1.0
stub = types.SimpleNamespace(GAMMA=0.99, LAMDA=0.98)
tr = [torch.tensor(r[:, :, k]) for k in range(4)]
adv = PPOAlgorithm.separate_cal_4_reward_GAE(stub, tr[0], tr[1], tr[2], tr[3], torch.tensor(v), torch.tensor(vn), torch.tensor(done))
adv = np.stack([a.numpy() for a in adv], axis=-1)
# local variant: streams (mk, pt, tt, it) take values jv[...,0], mv[...,0], mv[...,1], jv[...,1]
jv = torch.tensor(np.stack([v[:, :, 0], v[:, :, 3]], -1)); jvn = torch.tensor(np.stack([vn[:, :, 0], vn[:, :, 3]], -1))
mv = torch.tensor(np.stack([v[:, :, 1], v[:, :, 2]], -1)); mvn = torch.tensor(np.stack([vn[:, :, 1], vn[:, :, 2]], -1))
adv_l = PPOAlgorithm.cal_local_job_machine_reward_GAE(stub, tr[0], tr[1], tr[2], tr[3], jv, jvn, mv, mvn, torch.tensor(done))
adv_l = np.stack([a.numpy() for a in adv_l], axis=-1)
np.savez_compressed(os.path.join(HERE, "gae_golden.npz"), r=r, v=v, vn=vn, done=done, adv=adv, adv_local=adv_l)
print("wrote gae_golden.npz", adv.shape, float(np.abs(adv - adv_l).max()))
