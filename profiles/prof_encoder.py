"""Encoder timing / ncu driver: job-actor forward (GIN encoder + heads) on the native observation of workload A.
usage: python profiles/prof_encoder.py [envs] [precision] [reps]"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
precision = sys.argv[2] if len(sys.argv) > 2 else "tf32"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
pkg = importlib.import_module("e2e-mappo-for-mt-fjsp_b200")
envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
J, M, E = 6, 6, 2
d = pkg.instances.synthetic_instances(0, B, J, M, E, 1002)
env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
env.load(d["t"], d["p"], d["transT"], d["edge"])
env.scaler_init()
env.reset(pkg.instances.random_weights(0, B, 1002))
for s in range(12):
    env.random_step(seed=1)
job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(128), 11), J, M, precision=precision)
with torch.no_grad():
    for _ in range(2):
        out = job.forward(env.task_fea, env.adj_w, env.adj_src, env.candidate, None, env.job_mask, greedy=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = job.forward(env.task_fea, env.adj_w, env.adj_src, env.candidate, None, env.job_mask, greedy=True)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
rows = B * J * M
flops = 2.0 * rows * (12 * 128 + 5 * 128 * 128)
print("job actor forward: B=%d precision=%s  %.3f ms  (%.1f M env-forwards/s, encoder GEMMs %.1f TFLOP/s)"
      % (B, precision, ms, B / ms / 1e3, flops / ms / 1e9))
