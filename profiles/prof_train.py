"""One MAPPO training iteration on the device: collect `episodes` episodes with the tcgen05 rollout twins, then one
batched PPO update.  usage: python profiles/prof_train.py [envs] [episodes] [k_epochs] [mini_bs] [library_tf32 0|1] [encoder_tcgen05 0|1]"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
EP = int(sys.argv[2]) if len(sys.argv) > 2 else 1
KE = int(sys.argv[3]) if len(sys.argv) > 3 else 1
J, M, E, H = 6, 6, 2, 128
MB = int(sys.argv[4]) if len(sys.argv) > 4 else J * M
TF32 = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
ENC_TF32 = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
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
up = ppo.MAPPOUpdate(job, mch, crit, ppo.PPOConfig(k_epochs=KE, matmul_tf32=TF32, encoder_tf32=ENC_TF32))
ws = [pkg.instances.random_weights(0, B, 100 + e) for e in range(EP)]
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bt = ppo.collect(ro, ws)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    mean, std = up.update(bt, MB)
    ro.job.refresh(); ro.mch.refresh()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    steps = B * EP * J * M
    print("iter %d: collect %.1f ms  update %.1f ms  -> %.3e env-steps/s end to end; losses %s; peak mem %.1f GB"
          % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, steps / (t2 - t0), [round(float(x), 4) for x in mean],
             torch.cuda.max_memory_allocated() / 2 ** 30))
