"""Short driver for ncu captures: config-2 workload (J6M6E2, 65,536 envs), a few rollout steps.
usage: ncu ... python profiles/prof_step.py [workload] [steps] [random|replay|step]
  random: the one-launch random-rollout step; replay: mtfjsp_step_obs on given actions; step: mtfjsp_step alone"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "A"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
pkg = importlib.import_module("e2e-mappo-for-mt-fjsp_b200")
envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
J, M, E, B = wl["J"], wl["M"], wl["E"], wl["B"]
d = pkg.instances.synthetic_instances(0, B, J, M, E, wl["seed"])
w = pkg.instances.random_weights(0, B, wl["seed"])
env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
env.load(d["t"], d["p"], d["transT"], d["edge"])
env.scaler_init()
env.reset(w)
mode = sys.argv[3] if len(sys.argv) > 3 else "random"   # "random": one-launch random step; "replay": step_obs on given actions
if mode == "random":
    for s in range(steps):
        env.random_step(seed=1)
elif mode == "step":
    for s in range(steps):
        env.policy_random(seed=1)
        env.step(env.op, env.mach)
else:
    ops, mcs = [], []
    for s in range(steps):
        env.policy_random(seed=1)
        ops.append(env.op.clone()); mcs.append(env.mach.clone())
        env.step_obs(env.op, env.mach)
torch.cuda.synchronize()
print("done", env.launch_count)
