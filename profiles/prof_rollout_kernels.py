"""Kernel-time breakdown of actor-driven rollout steps (torch profiler, CUDA activities).
usage: python profiles/prof_rollout_kernels.py [envs] [steps]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
J, M, E = 6, 6, 2
pkg = importlib.import_module("e2e-mappo-for-mt-fjsp_b200")
envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
d = pkg.instances.synthetic_instances(0, B, J, M, E, 1002)
env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
env.load(d["t"], d["p"], d["transT"], d["edge"])
env.scaler_init()
job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(128), 11), J, M, precision="tf32")
mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(128), 12), M, precision="tf32")
ro = rom.Rollout(env, job, mch, greedy=False, use_cuda_graph=False, seed=1)
ro.begin_episode(pkg.instances.random_weights(0, B, 1002))
for _ in range(4):
    ro.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        ro.step()
    torch.cuda.synchronize()
print("per step = totals below / %d" % steps)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=80))
