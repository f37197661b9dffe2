"""Kernel-time breakdown of one PPO update (torch profiler, CUDA activities).  usage: python profiles/prof_train_kernels.py [envs]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ENC_TF32 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
J, M, E, H = 6, 6, 2, 128
pkg = importlib.import_module("e2e-mappo-for-mt-fjsp_b200")
envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
ppo = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.ppo")
d = pkg.instances.synthetic_instances(0, B, J, M, E, 1002)
env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
env.load(d["t"], d["p"], d["transT"], d["edge"])
env.scaler_init()
job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 1), J, M, hidden=H, trainable=True)
mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 2), M, hidden=H, trainable=True)
crit = enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(H), 3), J, M, hidden=H, trainable=True)
ro = rom.Rollout(env, job.inference_twin("tf32"), mch.inference_twin("tf32"), greedy=False, seed=3)
up = ppo.MAPPOUpdate(job, mch, crit, ppo.PPOConfig(k_epochs=1, encoder_tf32=ENC_TF32))
bt = ppo.collect(ro, [pkg.instances.random_weights(0, B, 100)])
up.update(bt, J * M)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    up.update(bt, J * M)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
